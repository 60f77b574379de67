// out = calz_vec_mex(cmd, ...)      plumbing of the calz_vec class (mex/calz_vec.m), handle mode of the gateways
//   h = calz_vec_mex('create', n, cols)          uint64 handle of a zeroed n x cols device block
//   calz_vec_mex('upload', h, col0, X)           X (n x k host double) -> columns [col0, col0+k)
//   X = calz_vec_mex('download', h, col0, cols)  columns [col0, col0+cols) as a host array
//   calz_vec_mex('copy', hdst, dcol0, hsrc, scol0, cols)   device-to-device column copy (Q(:,a:b) = Q_)
//   calz_vec_mex('free', h)
//   calz_vec_mex('clear_matrices')               drop the cached device matrices (explicit invalidation)
#include "calz_mex.h"
static calz_vec* handle_of(const mxArray* a) { return (calz_vec*)(uintptr_t)(*(const uint64_t*)mxGetData(a)); }
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[32] = {0};
    if (nrhs < 1 || !mxIsChar(prhs[0]) || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgIdAndTxt("calanczos:badarg", "calz_vec_mex(cmd, ...)");
    calz_ctx* ctx = calz_mex_context();
    if (!strcmp(cmd, "create") && nrhs == 3) {
        calz_vec* h = nullptr;
        calz_mex_fail(calz_vec_create(ctx, (int64_t)mxGetScalar(prhs[1]), (int)mxGetScalar(prhs[2]), &h), "calz_vec_create");
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)h;
    } else if (!strcmp(cmd, "upload") && nrhs == 4) {
        calz_mex_fail(calz_vec_upload(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), (int)mxGetN(prhs[3]), mxGetPr(prhs[3]),
                                      (int64_t)mxGetM(prhs[3])), "calz_vec_upload");
    } else if (!strcmp(cmd, "download") && nrhs == 4) {
        int64_t n = 0;
        calz_vec_info(handle_of(prhs[1]), nullptr, &n, nullptr, nullptr);
        const int cols = (int)mxGetScalar(prhs[3]);
        plhs[0] = mxCreateDoubleMatrix((size_t)n, (size_t)cols, mxREAL);
        calz_mex_fail(calz_vec_download(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), cols, mxGetPr(plhs[0]), n), "calz_vec_download");
    } else if (!strcmp(cmd, "copy") && nrhs == 6) {
        calz_mex_fail(calz_vec_copy(handle_of(prhs[1]), (int)mxGetScalar(prhs[2]), handle_of(prhs[3]), (int)mxGetScalar(prhs[4]),
                                    (int)mxGetScalar(prhs[5])), "calz_vec_copy");
    } else if (!strcmp(cmd, "free") && nrhs == 2) {
        calz_vec_destroy(handle_of(prhs[1]));
    } else if (!strcmp(cmd, "clear_matrices")) {
        calz_mat_cache_clear(ctx);
    } else {
        mexErrMsgIdAndTxt("calanczos:badarg", "calz_vec_mex: unknown command '%s'", cmd);
    }
    (void)nlhs;
}
