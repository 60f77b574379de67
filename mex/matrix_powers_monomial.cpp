// V = matrix_powers_monomial(A,q,s)               drop-in for matrix_powers_monomial.m:6-12  (n x s, q excluded)
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 3 || nlhs > 1) mexErrMsgIdAndTxt("calanczos:badarg", "usage: V = matrix_powers_monomial(A,q,s)");
    calz_mat* A = calz_mex_matrix(prhs[0]);
    const size_t n = mxGetM(prhs[0]);
    const int s = (int)mxGetScalar(prhs[2]);
    if (mxGetNumberOfElements(prhs[1]) != n || s < 1) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    plhs[0] = mxCreateDoubleMatrix(n, s, mxREAL);
    calz_mex_fail(calz_mpk_monomial_host(A, mxGetPr(prhs[1]), s, mxGetPr(plhs[0]), (int64_t)n), "matrix_powers_monomial");
}
