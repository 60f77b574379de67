// Av = SpMV(A,v)                                  drop-in for SpMV.m:6-8
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 2 || nlhs > 1) mexErrMsgIdAndTxt("calanczos:badarg", "usage: Av = SpMV(A,v)");
    calz_mat* A = calz_mex_matrix(prhs[0]);
    const size_t n = mxGetM(prhs[0]);
    if (mxGetNumberOfElements(prhs[1]) != n) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    plhs[0] = mxCreateDoubleMatrix(n, 1, mxREAL);
    calz_mex_fail(calz_spmv_host(A, mxGetPr(prhs[1]), mxGetPr(plhs[0])), "SpMV");
}
