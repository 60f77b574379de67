// [Q,R,rank] = normalize(X,opt,tol)               drop-in for normalize.m:3-36
// Handle mode: X a calz_vec => Q a calz_vec (R and rank are always host values).
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 1 || nrhs > 3 || nlhs > 3) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [Q,R,rank] = normalize(X,opt,tol)");
    if (nrhs >= 2 && mxIsChar(prhs[1])) {
        char opt[32] = {0};
        mxGetString(prhs[1], opt, sizeof(opt));
        if (!strcasecmp(opt, "randomizeNullSpace"))               // normalize.m:28-31, never requested by the drivers
            mexErrMsgIdAndTxt("calanczos:unsupported", "normalize(...,'randomizeNullSpace') is not on the hot path");
    }
    const double tol = (nrhs >= 3) ? mxGetScalar(prhs[2]) : 1.0e-8;            // :8-10
    calz_ctx* ctx = calz_mex_context();
    int rank = 0;
    mxArray *Q, *R;
    if (calz_mex_is_vec(prhs[0])) {
        CalzMexVec X = calz_mex_vec(prhs[0]), Qv;
        Q = calz_mex_new_vec(ctx, (size_t)X.n, X.cols, &Qv);
        R = mxCreateDoubleMatrix(X.cols, X.cols, mxREAL);
        calz_mex_fail(calz_normalize(ctx, X.n, X.cols, X.dev, X.ld, calz_mex_backend(), tol, Qv.dev, Qv.ld, mxGetPr(R), &rank), "normalize");
    } else {
        const size_t n = mxGetM(prhs[0]), c = mxGetN(prhs[0]);
        Q = mxCreateDoubleMatrix(n, c, mxREAL);
        R = mxCreateDoubleMatrix(c, c, mxREAL);
        calz_mex_fail(calz_normalize_host(ctx, (int64_t)n, (int)c, mxGetPr(prhs[0]), (int64_t)n, calz_mex_backend(), tol, mxGetPr(Q),
                                          (int64_t)n, mxGetPr(R), &rank), "normalize");
    }
    plhs[0] = Q;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar((double)rank);
}
