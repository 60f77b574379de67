// V = matrix_powers_monomial(A,q,s)               drop-in for matrix_powers_monomial.m:6-12  (n x s, q excluded)
// Handle mode: q a calz_vec => V a calz_vec.
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 3 || nlhs > 1) mexErrMsgIdAndTxt("calanczos:badarg", "usage: V = matrix_powers_monomial(A,q,s)");
    calz_mat* A = calz_mex_matrix(prhs[0]);
    const size_t n = mxGetM(prhs[0]);
    const int s = (int)mxGetScalar(prhs[2]);
    if (s < 1) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    if (calz_mex_is_vec(prhs[1])) {
        CalzMexVec q = calz_mex_vec(prhs[1]), V;
        if ((size_t)q.n != n || q.cols != 1) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
        plhs[0] = calz_mex_new_vec(calz_mex_context(), n, s, &V);
        calz_mex_fail(calz_mpk_monomial(A, q.dev, s, V.dev, V.ld), "matrix_powers_monomial");
        return;
    }
    if (mxGetNumberOfElements(prhs[1]) != n) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    plhs[0] = mxCreateDoubleMatrix(n, s, mxREAL);
    calz_mex_fail(calz_mpk_monomial_host(A, mxGetPr(prhs[1]), s, mxGetPr(plhs[0]), (int64_t)n), "matrix_powers_monomial");
}
