// Av = SpMV(A,v)                                  drop-in for SpMV.m:6-8
// Handle mode: v a calz_vec (n x 1) => Av a calz_vec.
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 2 || nlhs > 1) mexErrMsgIdAndTxt("calanczos:badarg", "usage: Av = SpMV(A,v)");
    calz_mat* A = calz_mex_matrix(prhs[0]);
    const size_t n = mxGetM(prhs[0]);
    if (calz_mex_is_vec(prhs[1])) {
        CalzMexVec v = calz_mex_vec(prhs[1]), y;
        if ((size_t)v.n != n || v.cols != 1) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
        plhs[0] = calz_mex_new_vec(calz_mex_context(), n, 1, &y);
        calz_mex_fail(calz_spmv(A, v.dev, y.dev), "SpMV");
        return;
    }
    if (mxGetNumberOfElements(prhs[1]) != n) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    plhs[0] = mxCreateDoubleMatrix(n, 1, mxREAL);
    calz_mex_fail(calz_spmv_host(A, mxGetPr(prhs[1]), mxGetPr(plhs[0])), "SpMV");
}
