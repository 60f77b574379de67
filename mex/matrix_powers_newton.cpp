// V = matrix_powers_newton(A,v,s,lambda,modifiedp)   drop-in for matrix_powers_newton.m:15-54  (n x (s+1))
// Handle mode: if v is a calz_vec (device block, mex/calz_vec.m) the basis comes back as a calz_vec too -- nothing crosses PCIe.
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 4 || nrhs > 5 || nlhs > 1)
        mexErrMsgIdAndTxt("calanczos:badarg", "usage: V = matrix_powers_newton(A,v,s,lambda,modifiedp)");
    calz_mat* A = calz_mex_matrix(prhs[0]);
    const size_t n = mxGetM(prhs[0]);
    const int s = (int)mxGetScalar(prhs[2]);
    const int modifiedp = (nrhs < 5) ? 0 : (int)mxGetScalar(prhs[4]);          // :16-18 default 0
    if (s < 1 || mxGetNumberOfElements(prhs[3]) < (size_t)s) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    const double* im = mxIsComplex(prhs[3]) ? mxGetPi(prhs[3]) : NULL;
    if (calz_mex_is_vec(prhs[1])) {
        CalzMexVec v = calz_mex_vec(prhs[1]), V;
        if ((size_t)v.n != n || v.cols != 1) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
        plhs[0] = calz_mex_new_vec(calz_mex_context(), n, s + 1, &V);
        calz_mex_fail(calz_mpk_newton(A, v.dev, s, mxGetPr(prhs[3]), im, modifiedp, V.dev, V.ld), "matrix_powers_newton");
        return;
    }
    if (mxGetNumberOfElements(prhs[1]) != n) mexErrMsgIdAndTxt("calanczos:badarg", "dimension mismatch");
    if (mxIsComplex(prhs[1])) mexErrMsgIdAndTxt("calanczos:unsupported", "complex start vectors are out of scope");
    plhs[0] = mxCreateDoubleMatrix(n, s + 1, mxREAL);
    calz_mex_fail(calz_mpk_newton_host(A, mxGetPr(prhs[1]), s, mxGetPr(prhs[3]), im, modifiedp, mxGetPr(plhs[0]), (int64_t)n),
                  "matrix_powers_newton");
}
