// [Q,R,rank] = normalize(X,opt,tol)               drop-in for normalize.m:3-36
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
    const size_t n = mxGetM(prhs[0]), c = mxGetN(prhs[0]);
    mxArray* Q = mxCreateDoubleMatrix(n, c, mxREAL);
    mxArray* R = mxCreateDoubleMatrix(c, c, mxREAL);
    int rank = 0;
    calz_mex_fail(calz_normalize_host(ctx, (int64_t)n, (int)c, mxGetPr(prhs[0]), (int64_t)n, calz_mex_backend(), tol, mxGetPr(Q),
                                      (int64_t)n, mxGetPr(R), &rank), "normalize");
    plhs[0] = Q;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
    if (nlhs > 2) plhs[2] = mxCreateDoubleScalar((double)rank);
}
