// [X,R] = project(Q,X,doreorth)                   drop-in for project.m:7-58  (Q cell array, R cell array)
// Handle mode: X (and every non-empty Q{i}) a calz_vec => the projected X comes back as a new calz_vec.
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 2 || nrhs > 3 || nlhs > 2) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [X,R] = project(Q,X,doreorth)");
    if (mxIsCell(prhs[1])) mexErrMsgIdAndTxt("calanczos:badarg", "Input X (arg 2) project() must be a column matrix.");   // :16-19
    const int doreorth = (nrhs >= 3) ? (mxIsLogicalScalarTrue(prhs[2]) || mxGetScalar(prhs[2]) != 0) : 0;                // :8-10
    calz_ctx* ctx = calz_mex_context();
    const bool dev = calz_mex_is_vec(prhs[1]);
    CalzMexVec Xin, Xout;
    if (dev) Xin = calz_mex_vec(prhs[1]);
    const size_t n = dev ? (size_t)Xin.n : mxGetM(prhs[1]), c = dev ? (size_t)Xin.cols : mxGetN(prhs[1]);
    CalzMexCell Q;
    calz_mex_cell(prhs[0], n, Q);
    const size_t nb = Q.ptr.size();
    for (size_t i = 0; i < nb; ++i)
        if (Q.mcols[i] > 0 && calz_mex_is_vec(mxGetCell(prhs[0], i)) != dev)
            mexErrMsgIdAndTxt("calanczos:badarg", "project: host arrays and calz_vec blocks cannot be mixed");
    // value semantics: never modify the input -- the projection runs in place on a copy (device-to-device in handle mode)
    mxArray* X;
    if (dev) {
        X = calz_mex_new_vec(ctx, n, (int)c, &Xout);
        calz_mex_fail(calz_vec_copy(Xout.h, 0, Xin.h, Xin.col0, (int)c), "project");
    } else {
        X = mxCreateDoubleMatrix(n, c, mxREAL);
        memcpy(mxGetPr(X), mxGetPr(prhs[1]), n * c * sizeof(double));
    }
    mxArray* R = mxCreateCellMatrix(1, nb);
    std::vector<double*> rp(nb, nullptr);
    for (size_t i = 0; i < nb; ++i)
        if (Q.mcols[i] > 0) {
            mxArray* Ri = mxCreateDoubleMatrix(Q.mcols[i], c, mxREAL);
            mxSetCell(R, i, Ri);
            rp[i] = mxGetPr(Ri);
        }
    int st = CALZ_OK;                                                           // :21-24 quick exit when there is no block
    if (nb && dev)
        st = calz_project(ctx, (int64_t)n, (int)nb, Q.ptr.data(), Q.ld.data(), Q.mcols.data(), (int)c, Xout.dev, Xout.ld, doreorth, rp.data());
    else if (nb)
        st = calz_project_host(ctx, (int64_t)n, (int)nb, Q.ptr.data(), Q.ld.data(), Q.mcols.data(), (int)c, mxGetPr(X), (int64_t)n,
                               doreorth, rp.data());
    { std::vector<double*>().swap(rp); CalzMexCell().ptr.swap(Q.ptr); }         // release before a possible longjmp
    calz_mex_fail(st, "project");
    plhs[0] = X;
    if (nlhs > 1) plhs[1] = R; else mxDestroyArray(R);
}
