// [Q,R] = cholqr(X)                               drop-in for cholqr.m:3-8  (errors like chol when X'X is not PD)
// Handle mode: X a calz_vec => Q a calz_vec (R is always a host matrix).
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs != 1 || nlhs > 2) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [Q,R] = cholqr(X)");
    calz_ctx* ctx = calz_mex_context();
    int info = 0;
    mxArray *Q, *R;
    if (calz_mex_is_vec(prhs[0])) {
        CalzMexVec X = calz_mex_vec(prhs[0]), Qv;
        Q = calz_mex_new_vec(ctx, (size_t)X.n, X.cols, &Qv);
        R = mxCreateDoubleMatrix(X.cols, X.cols, mxREAL);
        calz_mex_fail(calz_cholqr(ctx, X.n, X.cols, X.dev, X.ld, Qv.dev, Qv.ld, mxGetPr(R), &info), "cholqr");
    } else {
        const size_t n = mxGetM(prhs[0]), c = mxGetN(prhs[0]);
        Q = mxCreateDoubleMatrix(n, c, mxREAL);
        R = mxCreateDoubleMatrix(c, c, mxREAL);
        calz_mex_fail(calz_cholqr_host(ctx, (int64_t)n, (int)c, mxGetPr(prhs[0]), (int64_t)n, mxGetPr(Q), (int64_t)n, mxGetPr(R), &info),
                      "cholqr");
    }
    plhs[0] = Q;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
}
