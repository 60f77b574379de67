// Internal declarations shared by the libcalz translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/calz.h"

namespace calz {

// ---------------------------------------------------------------------------------------------------
// minimal NCCL surface, resolved with dlopen so that the library neither links against nor requires NCCL
// on single-GPU hosts (the MEX use case) and shares torch's copy when loaded into a torch process.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };   // ncclDataType_t: int8 0,uint8 1,int32 2,uint32 3,int64 4,uint64 5,f16 6,f32 7,f64 8
enum { ncclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
int nccl_load(const char* path, NcclApi** api, std::string* err);

// ---------------------------------------------------------------------------------------------------
struct DevBuf {                      // grow-only device scratch
    void* p = nullptr;
    size_t bytes = 0;
};

// ---------------------------------------------------------------------------------------------------
// peer-memory collectives (p2p.cu)
constexpr int kMaxPeers = 8;
constexpr int kMboxBufs = 4;             // mailbox generations (2 would do, see p2p.cu)
constexpr int kMboxSlot = 1024;          // doubles per (generation, source rank)
constexpr int kFlagAllreduce = 0;        // [kMboxBufs][kMaxPeers]
constexpr int kFlagHaloData = 32;        // [kMaxPeers] sequence number of the last halo pushed BY that peer
constexpr int kFlagHaloAck = 40;         // [kMaxPeers] sequence number of my last push that peer has consumed
constexpr int kFlagCount = 64;
struct P2P {
    bool enabled = false;
    void* base = nullptr;                // mailbox + flags, one IPC-exported allocation
    double* mbox = nullptr;
    unsigned long long* flags = nullptr;
    void* peer_base[kMaxPeers] = {};
    int* err = nullptr;                  // device flag: a bounded spin gave up
    unsigned int* tickets = nullptr;
    unsigned long long seq_allreduce = 0, seq_halo = 0;
    unsigned long long last_push[kMaxPeers] = {};
};

}  // namespace calz

struct calz_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;

    // communicator
    calz::NcclApi* nccl = nullptr;
    calz::ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    int setup_depth = 0;             // > 0 while a set-up phase runs: small all-reduces use NCCL, not the peer-memory mailbox

    // options
    int64_t opt_l2_chunk_bytes = 0;
    int64_t opt_sell_sigma = 0;      // 0: choose
    int64_t opt_csr_lanes = 0;       // 0: choose
    int64_t opt_cholqr2_inv_thresh = 32;
    int64_t opt_mpk_xs_rows = 0;     // cap on the rows per CTA of the TMA-staged kernel (0: 2048)
    int64_t opt_mpk_tma_x = 0;       // dictionary SELL: stage the x segments of a CTA in shared memory with TMA bulk copies
                                     // (measured at C3: 1.45 ms per MPK vs 1.19 ms for the L1-gather kernel => opt-in)
    int64_t opt_pan_fused_solve = 1; // tile pipeline, Cholesky back ends: downdated Gram + fused (update, triangular solve) last pass
    int64_t opt_mpk_dict_mode = -1;  // dictionary SELL: where the dictionary lives (0 shared memory, 2 constant bank, -1 by code uniformity)
    int64_t opt_mpk_halo_level = 0;  // depth of the ghost closure = MPK steps per halo exchange (0: automatic, see matrix.cu)
    int64_t opt_mpk_prefetch = 1;    // slice-pattern kernel: L2 prefetch distance of the leading edge of x, in warp iterations
                                     // (measured per C3 MPK: 0 -> 0.911 ms, 1 -> 0.765, 2 -> 0.773, 4 -> 0.779, 8 -> 0.790, 16 -> 0.911)
    int64_t opt_mpk_patterns = 1;    // dictionary SELL: slice-pattern kernel (k_spmv_selp) when most slices have a pattern
    int64_t opt_mpk_fused_steps = 0; // dictionary SELL: all steps of an exchange group in one cooperative launch (grid barriers); measured: no gain over
                                     // back-to-back launches (0.848 vs 0.850 ms per C3 MPK, 0.182 vs 0.176 ms on a 2.6 M-row slab) => off
    int64_t opt_mpk_persist = 1;     // dictionary SELL: persistent, software-pipelined kernel (0 = one CTA per 16 slices)
    int64_t opt_sell_dict = 1;       // layout=auto may pick the dictionary-coded SELL variant
    int64_t opt_fused_allreduce = 1; // tile passes: the finalize launch is the peer-memory all-reduce as well
    int64_t opt_p2p = 1;             // peer-memory all-reduce / halo push instead of NCCL (when IPC works)
    calz::P2P p2p;
    int64_t opt_mpk_pair_phase = -1; // slice pairing of the pattern kernel: -1 = the matrix' own choice, 0 / 1 forced (tests)
    int64_t opt_tile_panels = 1;     // tile kernels for multi-block / wide-block projections too, panel by panel (0: legacy kernels)
    int64_t opt_tile_pipeline = 1;   // fused TMA-tile passes in projectAndNormalize (0: legacy kernels)
    int64_t opt_grid_mult = 8;       // CTAs per SM for the persistent tall-skinny kernels

    // scratch
    calz::DevBuf partials;           // per-CTA partial Gram/coefficient tiles
    calz::DevBuf small;              // small device matrices (C, G, R, flags)
    calz::DevBuf work[4];            // n x c work blocks (Y, Z, host-flavour staging ...)
    calz::DevBuf tsqr_r;             // TSQR leaf R factors / tree
    void* tsqr_plan = nullptr;       // tsqr.cu: the live factorisation (levels, reflector storage); freed by tsqr_plan_free
    void* pan_ring = nullptr;        // orth.cu: ring of in-flight projectAndNormalize calls; freed by pan_ring_free
    unsigned int* ticket = nullptr;  // "last CTA" tickets
    double* pinned = nullptr;        // pinned host staging for small results
    size_t pinned_bytes = 0;
};

namespace calz {

extern std::string g_last_error;

int set_error(calz_ctx* ctx, int code, const char* fmt, ...);
int reserve(calz_ctx* ctx, DevBuf& b, size_t bytes);

#define CALZ_CUDA(ctx, expr)                                                                        \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return calz::set_error((ctx), CALZ_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, \
                                   cudaGetErrorString(_e));                                         \
    } while (0)

#define CALZ_TRY(expr)                 \
    do {                               \
        int _s = (expr);               \
        if (_s != CALZ_OK) return _s;  \
    } while (0)

#define CALZ_NCCL(ctx, expr)                                                                          \
    do {                                                                                              \
        int _e = (expr);                                                                              \
        if (_e != calz::ncclSuccess)                                                                  \
            return calz::set_error((ctx), CALZ_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #expr,   \
                                   (ctx)->nccl->GetErrorString(_e));                                  \
    } while (0)

#define CALZ_LAUNCH_CHECK(ctx)          \
    do {                                \
        (ctx)->launches++;              \
        CALZ_CUDA((ctx), cudaGetLastError()); \
    } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// in-place sum over ranks of a small device fp64 buffer (no-op without a communicator)
int allreduce_sum(calz_ctx* ctx, double* dev, size_t count);

// peer-memory collectives (p2p.cu)
int p2p_setup(calz_ctx* ctx);
void p2p_teardown(calz_ctx* ctx);
bool p2p_allreduce_ok(const calz_ctx* ctx, size_t count);
int p2p_allreduce(calz_ctx* ctx, double* dev, size_t count);
int p2p_check(calz_ctx* ctx);

void tsqr_plan_free(calz_ctx* ctx);      // tsqr.cu
void pan_ring_free(calz_ctx* ctx);       // orth.cu

// host small algebra (smallalg.cu)
void svd_singular_values(int c, const double* R, int ldR, double* sigma);   // one-sided Jacobi
int  numerical_rank(int c, const double* R, int ldR, double tol);            // normalize.m:15-24

}  // namespace calz
