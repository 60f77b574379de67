// Context, error handling, scratch management, NCCL plumbing and host small algebra of libcalz.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "matrix.h"

namespace calz {

std::string g_last_error;

int set_error(calz_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->err = buf;
    return code;
}

int reserve(calz_ctx* ctx, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return CALZ_OK;
    if (b.p) {
        // the old block may still be in use by kernels queued on the stream
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CALZ_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return set_error(ctx, CALZ_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    b.bytes = want;
    return CALZ_OK;
}

// --------------------------------------------------------------------------------------------- NCCL
static NcclApi g_nccl;

int nccl_load(const char* path, NcclApi** api, std::string* err) {
    if (g_nccl.handle) {
        *api = &g_nccl;
        return CALZ_OK;
    }
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // torch's copy, if resident
    if (!h) {
        const char* env = getenv("CALZ_NCCL_LIB");
        if (env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        *err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return CALZ_ERR_NCCL;
    }
    NcclApi a;
    a.handle = h;
#define CALZ_SYM(field, name)                                    \
    *(void**)(&a.field) = dlsym(h, name);                        \
    if (!a.field) {                                              \
        *err = std::string("libnccl: missing symbol ") + name;   \
        return CALZ_ERR_NCCL;                                    \
    }
    CALZ_SYM(GetUniqueId, "ncclGetUniqueId")
    CALZ_SYM(CommInitRank, "ncclCommInitRank")
    CALZ_SYM(CommDestroy, "ncclCommDestroy")
    CALZ_SYM(AllReduce, "ncclAllReduce")
    CALZ_SYM(Send, "ncclSend")
    CALZ_SYM(Recv, "ncclRecv")
    CALZ_SYM(GroupStart, "ncclGroupStart")
    CALZ_SYM(GroupEnd, "ncclGroupEnd")
    CALZ_SYM(GetErrorString, "ncclGetErrorString")
#undef CALZ_SYM
    g_nccl = a;
    *api = &g_nccl;
    return CALZ_OK;
}

int allreduce_sum(calz_ctx* ctx, double* dev, size_t count) {
    if (ctx->nranks <= 1 || count == 0) return CALZ_OK;
    // set-up phases (matrix creation) go through NCCL: ranks may arrive tens of seconds apart there (host-side matrix generation),
    // which a stream-ordered NCCL collective simply waits out while the bounded spins of the peer-memory kernels would time out
    if (ctx->setup_depth == 0 && p2p_allreduce_ok(ctx, count)) return p2p_allreduce(ctx, dev, count);
    CALZ_NCCL(ctx, ctx->nccl->AllReduce(dev, dev, count, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    return CALZ_OK;
}

// --------------------------------------------------------------------------------------------- small algebra
// Singular values of a small c x c matrix by one-sided (Hestenes) Jacobi; replaces svd(R) in normalize.m:15.
void svd_singular_values(int c, const double* R, int ldR, double* sigma) {
    std::vector<double> U((size_t)c * c);
    for (int j = 0; j < c; ++j)
        for (int i = 0; i < c; ++i) U[(size_t)j * c + i] = R[(size_t)j * ldR + i];
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < c - 1; ++p)
            for (int q = p + 1; q < c; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < c; ++i) {
                    double up = U[(size_t)p * c + i], uq = U[(size_t)q * c + i];
                    a += up * up; b += uq * uq; g += up * uq;
                }
                if (g == 0.0) continue;
                double denom = sqrt(a * b);
                if (denom > 0 && fabs(g) <= 1e-17 * denom) continue;
                off = std::max(off, denom > 0 ? fabs(g) / denom : 0.0);
                double zeta = (b - a) / (2.0 * g);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < c; ++i) {
                    double up = U[(size_t)p * c + i], uq = U[(size_t)q * c + i];
                    U[(size_t)p * c + i] = cs * up - sn * uq;
                    U[(size_t)q * c + i] = sn * up + cs * uq;
                }
            }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < c; ++j) {
        double a = 0;
        for (int i = 0; i < c; ++i) a += U[(size_t)j * c + i] * U[(size_t)j * c + i];
        sigma[j] = sqrt(a);
    }
    std::sort(sigma, sigma + c, [](double x, double y) { return x > y; });
}

// normalize.m:15-24: rank = index of the first sigma_i <= tol*sigma_1, minus one; ncols if none.
int numerical_rank(int c, const double* R, int ldR, double tol) {
    std::vector<double> s(c);
    svd_singular_values(c, R, ldR, s.data());
    double abs_tol = tol * s[0];
    for (int i = 0; i < c; ++i)
        if (!(s[i] > abs_tol)) return i;
    return c;
}

}  // namespace calz

using namespace calz;

// --------------------------------------------------------------------------------------------- C ABI
extern "C" {

int calz_version(void) { return 100; }

const char* calz_last_error(const calz_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int calz_init(int device, calz_ctx** out) {
    if (!out) return set_error(nullptr, CALZ_ERR_BADARG, "calz_init: ctx is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_error(nullptr, CALZ_ERR_CUDA, "calz_init: no CUDA device (%s); there is no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return set_error(nullptr, CALZ_ERR_BADARG, "calz_init: bad device %d", device);
    calz_ctx* ctx = new calz_ctx();
    ctx->device = device;
    CALZ_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    CALZ_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete ctx;
        return set_error(nullptr, CALZ_ERR_CUDA, "calz_init: device sm_%d%d is not a Blackwell B200 (sm_100a) part",
                         prop.major, prop.minor);
    }
    ctx->num_sms = prop.multiProcessorCount;
    CALZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    CALZ_CUDA(ctx, cudaMalloc(&ctx->ticket, 64 * sizeof(unsigned int)));
    CALZ_CUDA(ctx, cudaMemset(ctx->ticket, 0, 64 * sizeof(unsigned int)));
    ctx->pinned_bytes = 1 << 20;
    CALZ_CUDA(ctx, cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    *out = ctx;
    return CALZ_OK;
}

int calz_finalize(calz_ctx* ctx) {
    if (!ctx) return CALZ_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    p2p_teardown(ctx);
    tsqr_plan_free(ctx);
    pan_ring_free(ctx);
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
    DevBuf* bufs[] = {&ctx->partials, &ctx->small, &ctx->work[0], &ctx->work[1], &ctx->work[2], &ctx->work[3], &ctx->tsqr_r};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->ticket) cudaFree(ctx->ticket);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return CALZ_OK;
}

int calz_set_stream(calz_ctx* ctx, void* s) {
    if (!ctx) return CALZ_ERR_BADARG;
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (s) {
        ctx->stream = (cudaStream_t)s;
        ctx->own_stream = false;
    } else {
        CALZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return CALZ_OK;
}

void* calz_get_stream(calz_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int calz_sync(calz_ctx* ctx) {
    if (!ctx) return CALZ_ERR_BADARG;
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return p2p_check(ctx);
}

int64_t calz_launch_count(calz_ctx* ctx, int reset) {
    if (!ctx) return 0;
    int64_t v = ctx->launches;
    if (reset) ctx->launches = 0;
    return v;
}

int calz_set_option(calz_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return CALZ_ERR_BADARG;
    if (!strcmp(key, "mpk_l2_chunk_bytes")) ctx->opt_l2_chunk_bytes = value;
    else if (!strcmp(key, "sell_sigma")) ctx->opt_sell_sigma = value;
    else if (!strcmp(key, "csr_lanes")) ctx->opt_csr_lanes = value;
    else if (!strcmp(key, "cholqr2_inv_thresh")) ctx->opt_cholqr2_inv_thresh = value > 0 ? value : 32;
    else if (!strcmp(key, "p2p")) ctx->opt_p2p = value;
    else if (!strcmp(key, "fused_allreduce")) ctx->opt_fused_allreduce = value;
    else if (!strcmp(key, "sell_dict")) ctx->opt_sell_dict = value;
    else if (!strcmp(key, "mpk_tma_x")) ctx->opt_mpk_tma_x = value;
    else if (!strcmp(key, "pan_fused_solve")) ctx->opt_pan_fused_solve = value;
    else if (!strcmp(key, "mpk_persist")) ctx->opt_mpk_persist = value;
    else if (!strcmp(key, "mpk_halo_level")) ctx->opt_mpk_halo_level = value;
    else if (!strcmp(key, "mpk_fused_steps")) ctx->opt_mpk_fused_steps = value;
    else if (!strcmp(key, "mpk_patterns")) ctx->opt_mpk_patterns = value;
    else if (!strcmp(key, "mpk_prefetch")) ctx->opt_mpk_prefetch = value;
    else if (!strcmp(key, "mpk_dict_mode")) ctx->opt_mpk_dict_mode = value;
    else if (!strcmp(key, "mpk_xs_rows")) ctx->opt_mpk_xs_rows = value;
    else if (!strcmp(key, "tile_pipeline")) ctx->opt_tile_pipeline = value;
    else if (!strcmp(key, "tile_panels")) ctx->opt_tile_panels = value;
    else if (!strcmp(key, "mpk_pair_phase")) ctx->opt_mpk_pair_phase = value;
    else if (!strcmp(key, "grid_mult")) ctx->opt_grid_mult = value > 0 ? value : 1;
    else return set_error(ctx, CALZ_ERR_BADARG, "calz_set_option: unknown key '%s'", key);
    return CALZ_OK;
}

// ---- process-wide context + device-matrix cache (the MEX gateways' shared state)
namespace {
struct CachedMat { const void* pr; int64_t n; uint64_t nnz, fp; int s_max, layout; calz_mat* m; };
calz_ctx* g_shared_ctx = nullptr;
std::vector<CachedMat> g_mat_cache;

uint64_t fnv(uint64_t h, const void* p, size_t bytes) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
// 4096 strided samples of each array plus 64 entries at both ends: O(1) per call, catches a rebuilt matrix at a reused address
uint64_t csc_fingerprint(int64_t n, uint64_t nnz, const uint64_t* jc, const uint64_t* ir, const double* pr) {
    uint64_t h = 1469598103934665603ull;
    auto sample = [&](const void* base, size_t count, size_t elem) {
        if (!count) return;
        const size_t step = std::max<size_t>(1, count / 4096);
        for (size_t i = 0; i < count; i += step) h = fnv(h, (const char*)base + i * elem, elem);
        const size_t edge = std::min<size_t>(64, count);
        h = fnv(h, base, edge * elem);
        h = fnv(h, (const char*)base + (count - edge) * elem, edge * elem);
    };
    sample(jc, (size_t)n + 1, 8);
    sample(ir, (size_t)nnz, 8);
    sample(pr, (size_t)nnz, 8);
    return h;
}
}  // namespace

int calz_shared_context(calz_ctx** ctx) {
    if (!ctx) return set_error(nullptr, CALZ_ERR_BADARG, "calz_shared_context: ctx is NULL");
    if (!g_shared_ctx) {
        const char* dev = getenv("CALZ_DEVICE");
        CALZ_TRY(calz_init(dev ? atoi(dev) : 0, &g_shared_ctx));
    }
    *ctx = g_shared_ctx;
    return CALZ_OK;
}

int calz_mat_cache_clear(calz_ctx* ctx) {
    for (size_t i = 0; i < g_mat_cache.size();) {
        if (!ctx || g_mat_cache[i].m->ctx == ctx) {
            calz_mat_destroy(g_mat_cache[i].m);
            g_mat_cache.erase(g_mat_cache.begin() + i);
        } else {
            ++i;
        }
    }
    return CALZ_OK;
}

int calz_shared_release(void) {
    calz_mat_cache_clear(nullptr);
    int st = CALZ_OK;
    if (g_shared_ctx) st = calz_finalize(g_shared_ctx);
    g_shared_ctx = nullptr;
    return st;
}

int calz_mat_cache_get_csc64(calz_ctx* ctx, int64_t n, const uint64_t* jc, const uint64_t* ir, const double* pr, int s_max,
                             int layout, calz_mat** mat) {
    if (!ctx || !jc || !ir || !pr || !mat || n <= 0) return set_error(ctx, CALZ_ERR_BADARG, "calz_mat_cache_get_csc64: bad arguments");
    const uint64_t nnz = jc[n];
    const uint64_t fp = csc_fingerprint(n, nnz, jc, ir, pr);
    for (size_t i = 0; i < g_mat_cache.size(); ++i) {
        CachedMat& c = g_mat_cache[i];
        if (c.m->ctx != ctx || c.pr != (const void*)pr || c.n != n || c.nnz != nnz) continue;
        if (c.fp == fp && c.s_max >= s_max && c.layout == layout) { *mat = c.m; return CALZ_OK; }
        calz_mat_destroy(c.m);                                  // same address, different content (or too small an s_max): stale
        g_mat_cache.erase(g_mat_cache.begin() + i);
        break;
    }
    if (g_mat_cache.size() >= 4) calz_mat_cache_clear(ctx);     // small cache: drop everything when it fills up
    calz_mat* m = nullptr;
    CALZ_TRY(calz_mat_create_csc64(ctx, n, jc, ir, pr, s_max, layout, &m));
    g_mat_cache.push_back(CachedMat{(const void*)pr, n, nnz, fp, s_max, layout, m});
    *mat = m;
    return CALZ_OK;
}

// ---- device blocks
}  // extern "C"
struct calz_vec {
    calz_ctx* ctx;
    double* d;
    int64_t n, ld;
    int cols;
};
extern "C" {

int calz_vec_create(calz_ctx* ctx, int64_t n, int cols, calz_vec** out) {
    if (!ctx || !out || n < 1 || cols < 1) return set_error(ctx, CALZ_ERR_BADARG, "calz_vec_create: bad arguments");
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    calz_vec* v = new calz_vec{ctx, nullptr, n, round_up(n, 32), cols};
    const size_t bytes = (size_t)v->ld * cols * sizeof(double);
    cudaError_t e = cudaMalloc(&v->d, bytes);
    if (e != cudaSuccess) {
        delete v;
        return set_error(ctx, CALZ_ERR_ALLOC, "calz_vec_create: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    CALZ_CUDA(ctx, cudaMemsetAsync(v->d, 0, bytes, ctx->stream));
    *out = v;
    return CALZ_OK;
}

int calz_vec_destroy(calz_vec* v) {
    if (!v) return CALZ_OK;
    cudaStreamSynchronize(v->ctx->stream);
    cudaFree(v->d);
    delete v;
    return CALZ_OK;
}

int calz_vec_info(const calz_vec* v, double** dev, int64_t* n, int* cols, int64_t* ld) {
    if (!v) return CALZ_ERR_BADARG;
    if (dev) *dev = v->d;
    if (n) *n = v->n;
    if (cols) *cols = v->cols;
    if (ld) *ld = v->ld;
    return CALZ_OK;
}

int calz_vec_upload(calz_vec* v, int col0, int cols, const double* host, int64_t ldh) {
    if (!v || !host || col0 < 0 || cols < 1 || col0 + cols > v->cols || ldh < v->n)
        return set_error(v ? v->ctx : nullptr, CALZ_ERR_BADARG, "calz_vec_upload: bad arguments");
    calz_ctx* ctx = v->ctx;
    CALZ_CUDA(ctx, cudaMemcpy2DAsync(v->d + (size_t)col0 * v->ld, (size_t)v->ld * sizeof(double), host, (size_t)ldh * sizeof(double),
                                     (size_t)v->n * sizeof(double), (size_t)cols, cudaMemcpyHostToDevice, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // the host buffer may be reused at once (MATLAB value semantics)
    return CALZ_OK;
}

int calz_vec_download(const calz_vec* v, int col0, int cols, double* host, int64_t ldh) {
    if (!v || !host || col0 < 0 || cols < 1 || col0 + cols > v->cols || ldh < v->n)
        return set_error(v ? v->ctx : nullptr, CALZ_ERR_BADARG, "calz_vec_download: bad arguments");
    calz_ctx* ctx = v->ctx;
    CALZ_CUDA(ctx, cudaMemcpy2DAsync(host, (size_t)ldh * sizeof(double), v->d + (size_t)col0 * v->ld, (size_t)v->ld * sizeof(double),
                                     (size_t)v->n * sizeof(double), (size_t)cols, cudaMemcpyDeviceToHost, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CALZ_OK;
}

int calz_vec_copy(calz_vec* dst, int dcol0, const calz_vec* src, int scol0, int cols) {
    if (!dst || !src || dst->n != src->n || cols < 1 || dcol0 < 0 || scol0 < 0 || dcol0 + cols > dst->cols || scol0 + cols > src->cols)
        return set_error(dst ? dst->ctx : nullptr, CALZ_ERR_BADARG, "calz_vec_copy: bad arguments");
    calz_ctx* ctx = dst->ctx;
    CALZ_CUDA(ctx, cudaMemcpy2DAsync(dst->d + (size_t)dcol0 * dst->ld, (size_t)dst->ld * sizeof(double), src->d + (size_t)scol0 * src->ld,
                                     (size_t)src->ld * sizeof(double), (size_t)src->n * sizeof(double), (size_t)cols,
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    return CALZ_OK;
}

int calz_comm_unique_id(char id_out[128], const char* nccl_lib) {
    NcclApi* api = nullptr;
    std::string err;
    if (nccl_load(nccl_lib, &api, &err) != CALZ_OK) return set_error(nullptr, CALZ_ERR_NCCL, "%s", err.c_str());
    ncclUniqueId id;
    int e = api->GetUniqueId(&id);
    if (e != ncclSuccess) return set_error(nullptr, CALZ_ERR_NCCL, "ncclGetUniqueId: %s", api->GetErrorString(e));
    memcpy(id_out, id.internal, 128);
    return CALZ_OK;
}

int calz_comm_init(calz_ctx* ctx, int nranks, int rank, const char id[128], const char* nccl_lib) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return set_error(ctx, CALZ_ERR_BADARG, "calz_comm_init: bad arguments");
    if (nranks == 1) {
        ctx->rank = 0;
        ctx->nranks = 1;
        return CALZ_OK;
    }
    std::string err;
    if (nccl_load(nccl_lib, &ctx->nccl, &err) != CALZ_OK) return set_error(ctx, CALZ_ERR_NCCL, "%s", err.c_str());
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    CALZ_NCCL(ctx, ctx->nccl->CommInitRank(&ctx->comm, nranks, uid, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
    return p2p_setup(ctx);        // peer-memory mailbox over CUDA IPC; silently stays on NCCL if IPC is unavailable
}

int calz_comm_rank(const calz_ctx* ctx, int* rank, int* nranks) {
    if (!ctx) return CALZ_ERR_BADARG;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return CALZ_OK;
}

}  // extern "C"
