// [QZ,RZ] = projectAndNormalize(Q,X,doreorth)     drop-in for projectAndNormalize.m:3-90
// RZ is a 1 x (numBlocks+1) cell: pass-1 + pass-2 coefficients per block (:71-73), RZ{end} = R of the last normalize.
// Handle mode: X (and every non-empty Q{i}) a calz_vec => QZ a calz_vec; the small RZ blocks are host arrays as always.
#include "calz_mex.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    if (nrhs < 2 || nrhs > 3 || nlhs > 2) mexErrMsgIdAndTxt("calanczos:badarg", "usage: [QZ,RZ] = projectAndNormalize(Q,X,doreorth)");
    const int doreorth = (nrhs >= 3) ? (mxIsLogicalScalarTrue(prhs[2]) || mxGetScalar(prhs[2]) != 0) : 1;                // :5-7
    calz_ctx* ctx = calz_mex_context();
    const bool dev = calz_mex_is_vec(prhs[1]);
    CalzMexVec Xv, QZv;
    if (dev) Xv = calz_mex_vec(prhs[1]);
    const size_t n = dev ? (size_t)Xv.n : mxGetM(prhs[1]), c = dev ? (size_t)Xv.cols : mxGetN(prhs[1]);
    CalzMexCell Q;
    calz_mex_cell(prhs[0], n, Q);
    const size_t nb = Q.ptr.size();
    for (size_t i = 0; i < nb; ++i)
        if (Q.mcols[i] > 0 && calz_mex_is_vec(mxGetCell(prhs[0], i)) != dev)
            mexErrMsgIdAndTxt("calanczos:badarg", "projectAndNormalize: host arrays and calz_vec blocks cannot be mixed");
    mxArray* QZ = dev ? calz_mex_new_vec(ctx, n, (int)c, &QZv) : mxCreateDoubleMatrix(n, c, mxREAL);
    mxArray* RZ = mxCreateCellMatrix(1, nb + 1);
    std::vector<double*> rp(nb + 1, nullptr);
    for (size_t i = 0; i < nb; ++i)
        if (Q.mcols[i] > 0) {
            mxArray* Ri = mxCreateDoubleMatrix(Q.mcols[i], c, mxREAL);
            mxSetCell(RZ, i, Ri);
            rp[i] = mxGetPr(Ri);
        }
    mxArray* Rl = mxCreateDoubleMatrix(c, c, mxREAL);
    mxSetCell(RZ, nb, Rl);
    int second = 0, rank = 0;
    int st;
    if (dev)
        st = calz_project_and_normalize(ctx, (int64_t)n, (int)nb, Q.ptr.data(), Q.ld.data(), Q.mcols.data(), (int)c, Xv.dev, Xv.ld,
                                        doreorth, calz_mex_backend(), QZv.dev, QZv.ld, rp.data(), mxGetPr(Rl), &second, &rank);
    else
        st = calz_project_and_normalize_host(ctx, (int64_t)n, (int)nb, Q.ptr.data(), Q.ld.data(), Q.mcols.data(), (int)c,
                                             mxGetPr(prhs[1]), (int64_t)n, doreorth, calz_mex_backend(), mxGetPr(QZ), (int64_t)n,
                                             rp.data(), mxGetPr(Rl), &second, &rank);
    { std::vector<double*>().swap(rp); CalzMexCell().ptr.swap(Q.ptr); }
    calz_mex_fail(st, "projectAndNormalize");
    if (second) mexPrintf("second\n");                                          // projectAndNormalize.m:62
    if (rank < (int)c && second) mexPrintf("Rank deficient\n");                 // :79-80
    plhs[0] = QZ;
    if (nlhs > 1) plhs[1] = RZ; else mxDestroyArray(RZ);
}
