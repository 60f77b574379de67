// Fused projectAndNormalize passes on TMA-staged row tiles (the steady-state block of ca_lanczos_basic).
//
// One persistent CTA per resident slot streams 128-row tiles of the tall-skinny operands [Q | X] through a
// multi-stage shared-memory ring: a producer warp issues one 1-D TMA bulk copy (cp.async.bulk, mbarrier
// complete_tx) per column, 8 consumer warps
//   - (update modes) apply  Y = X - Q*C  with fp64 DMMA on the 16 rows the warp owns, in shared memory, and stream Y back to HBM,
//   - contract the tile over its rows with fp64 DMMA (mma.sync m8n8k4), fragments read conflict-free from
//     shared memory (column pitch 132 doubles),
// so every operand crosses HBM exactly once per pass:
//   pass 1  COEFF   S1 = [Q X]'X               -> C1 = Q'X (project.m:34) and diag(X'X) = ||x_i||^2 (pAN.m:17-22)
//   pass 2  UPDATE  Y = X - Q*C1 (project.m:35), S2 = [Q Y]'Y -> C2 = Q'Y (2nd pass, pAN.m:63) and G_Y = Y'Y (cholqr.m:5)
//   pass 3  UPDATE  Z = Y - Q*C2 in place,      S3 = Z'Z       (cholqr.m:5 of the second normalize, pAN.m:64)
//   pass 3' SOLVE   QZ = ((X - Q*C1) - Q*C2) / R, nothing contracted (R from the downdated Gram, k_chol_pan): replaces
//                   pass 3 AND the separate triangular solve for the Cholesky back ends.  Y = X - Q*C1 is RE-computed in
//                   registers (bit-identical to pass 2), so pass 2 does not have to write Y to HBM at all.
// Reductions are deterministic: fixed warp order inside the CTA, per-CTA partials, last CTA sums them in a fixed order.
#include <algorithm>

#include "p2p_dev.cuh"
#include "tsops.cuh"

namespace calz {

namespace {

constexpr int kTileRows = 128;
constexpr int kPitch = 132;                     // doubles per column slot: 132 mod 16 == 4 => conflict-free DMMA fragments
constexpr int kConsumerWarps = 8;
constexpr int kTileThreads = (kConsumerWarps + 1) * 32;

enum { MODE_COEFF = 0, MODE_UPDATE_FULL = 1, MODE_UPDATE_GRAM = 2, MODE_UPDATE_SOLVE = 3 };

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct TileArgs {
    long long n;
    const double* Q; long long ldQ; int M;          // A panel (previous block)
    const double* X; long long ldX; int c;          // block being orthogonalised (input of this pass)
    double* Y; long long ldY;                       // update modes: where the updated block goes (may alias X)
    const double* C; int ldC;                       // update modes: coefficients (M x c, device)
    const double* R;                                // solve mode: c x c upper triangular factor (device, dense)
    const double* C2; int ldC2; const int* flag2;   // solve mode: second coefficient block, applied iff *flag2 != 0
    double* S; int ldS;                             // result: rows [0,M) = Q-part, rows [M, M+c) = block-part  (x c columns)
    double* partials; unsigned int* ticket;
    const int* pred; int want;
    int use_tma;                                    // 0: operands not 16-byte aligned on this rank -> plain loads (same arithmetic)
};

// dynamic shared memory layout: [stages][slots][kPitch] doubles | Cs[M][8*CT] | red | barriers
template <int MT, int CT, int MODE>
__global__ void __launch_bounds__(kTileThreads, 1)
k_tile(TileArgs p, int stages) {
    if (p.pred && *p.pred != p.want) return;
    constexpr int SLOTS = 8 * (MT + CT);
    constexpr bool SOLVE = (MODE == MODE_UPDATE_SOLVE);
    constexpr bool ACC_Q = (MODE != MODE_UPDATE_GRAM) && !SOLVE;  // accumulate the Q-row tiles too
    constexpr int RT0 = ACC_Q ? 0 : MT;                           // first row-tile that is accumulated
    constexpr int NRT = MT + CT - RT0;
    constexpr int CW = 8 * CT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    const size_t stage_doubles = (size_t)SLOTS * kPitch;
    double* Cs = tiles + (size_t)stages * stage_doubles;          // [MT*8][CW]
    double* red = Cs + (size_t)MT * 8 * CW;                       // [kConsumerWarps][NRT*CT*64]   (solve mode: Rs[CW][CW+1], Cs2)
    double* Cs2 = red + (size_t)CW * (CW + 1);                    // solve mode only: [MT*8][CW]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + (SOLVE ? (size_t)CW * (CW + 1) + (size_t)MT * 8 * CW + CW : (size_t)kConsumerWarps * NRT * CT * 64));
    uint64_t* empty = full + stages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (p.n + kTileRows - 1) / kTileRows;

    // ---- one-time set-up: barriers, zero the padding slots, stage C
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
        fence_barrier_init();
    }
    {   // only the padding slots (columns M..8MT-1 and c..8CT-1 of every stage) are ever read without being written by TMA
        const int padQ = 8 * MT - p.M, npad = padQ + 8 * CT - p.c;
        for (int e = tid; e < stages * npad * kTileRows; e += kTileThreads) {
            const int r = e % kTileRows, k = (e / kTileRows) % npad, st = e / (kTileRows * npad);
            const int slot = k < padQ ? p.M + k : 8 * MT + p.c + (k - padQ);
            tiles[(size_t)st * stage_doubles + (size_t)slot * kPitch + r] = 0.0;
        }
    }
    if (MODE != MODE_COEFF)
        for (int e = tid; e < MT * 8 * CW; e += kTileThreads) {
            const int m = e / CW, j = e % CW;
            Cs[e] = (m < p.M && j < p.c) ? p.C[(size_t)j * p.ldC + m] : 0.0;
        }
    const bool second = SOLVE && p.flag2 && *p.flag2 != 0;
    if (SOLVE) {
        for (int e = tid; e < CW * CW; e += kTileThreads) {
            const int i = e / CW, j = e % CW;
            red[i * (CW + 1) + j] = (i < p.c && j < p.c) ? p.R[(size_t)j * p.c + i] : (i == j ? 1.0 : 0.0);
        }
        for (int e = tid; e < MT * 8 * CW; e += kTileThreads) {
            const int m = e / CW, j = e % CW;
            Cs2[e] = (second && m < p.M && j < p.c) ? p.C2[(size_t)j * p.ldC2 + m] : 0.0;
        }
        for (int j = tid; j < CW; j += kTileThreads) Cs2[(size_t)MT * 8 * CW + j] = j < p.c ? 1.0 / p.R[(size_t)j * p.c + j] : 1.0;
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // =========================================================== producer warp
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % stages;
            const uint32_t ph = (uint32_t)((it / stages) & 1);
            mbar_wait(&empty[s], ph ^ 1);
            double* dst = tiles + (size_t)s * stage_doubles;
            const long long r0 = t * kTileRows;
            const int valid = (int)min((long long)kTileRows, p.n - r0);
            if (valid == kTileRows && p.use_tma) {
                if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)((p.M + p.c) * kTileRows * sizeof(double)));
                __syncwarp();
                for (int col = lane; col < p.M + p.c; col += 32) {
                    const bool isq = col < p.M;
                    const double* src = isq ? p.Q + (long long)col * p.ldQ + r0 : p.X + (long long)(col - p.M) * p.ldX + r0;
                    const int slot = isq ? col : 8 * MT + (col - p.M);
                    tma_bulk_g2s(dst + (size_t)slot * kPitch, src, kTileRows * sizeof(double), &full[s]);
                }
            } else {
                // ragged last tile (or operands that are not 16-byte aligned): plain loads, zero fill, then a plain arrive
                for (int e = lane; e < (p.M + p.c) * kTileRows; e += 32) {
                    const int col = e / kTileRows, r = e % kTileRows;
                    const bool isq = col < p.M;
                    const int slot = isq ? col : 8 * MT + (col - p.M);
                    double v = 0.0;
                    if (r < valid) v = isq ? p.Q[(long long)col * p.ldQ + r0 + r] : p.X[(long long)(col - p.M) * p.ldX + r0 + r];
                    dst[(size_t)slot * kPitch + r] = v;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
    } else {
        // =========================================================== consumer warps
        const int g = lane >> 2, tq = lane & 3;
        double acc[NRT][CT][2];
#pragma unroll
        for (int a = 0; a < NRT; ++a)
#pragma unroll
            for (int b = 0; b < CT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        // A fragments of the update GEMMs: -C1' (and -C2' for the solve pass), constant over the tiles
        double a1[2 * MT][CT], a2[2 * MT][CT];
#pragma unroll
        for (int ks = 0; ks < 2 * MT; ++ks)
#pragma unroll
            for (int b = 0; b < CT; ++b) {
                a1[ks][b] = (MODE != MODE_COEFF) ? -Cs[(4 * ks + tq) * CW + 8 * b + g] : 0.0;
                a2[ks][b] = SOLVE ? -Cs2[(4 * ks + tq) * CW + 8 * b + g] : 0.0;
            }
        int it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const int s = it % stages;
            const uint32_t ph = (uint32_t)((it / stages) & 1);
            mbar_wait(&full[s], ph);
            double* T = tiles + (size_t)s * stage_doubles;
            const long long r0 = t * kTileRows;
            if (MODE != MODE_COEFF) {
                // ---- Y = X - Q*C on the tensor pipe, warp-local: the warp updates the SAME 16 rows it contracts below, so no
                //      CTA-wide barrier is needed.  Per 8-row group: D(8 cols x 8 rows) = X' + (-C')(8 x K) * Q'(K x 8 rows);
                //      thread (g, tq) holds D[col g][n = 2tq, 2tq+1], A[col g][k tq] = -C(k, col), B[k tq][n g] = Q(row, k).
                //      The n index of the fragment is mapped to the tile row rho(n) = {0,1,2,3,5,4,7,6}[n]: with the 132-double
                //      column pitch this makes the D-fragment accesses (two 8-byte words per thread, rows rho(2tq), rho(2tq+1))
                //      AND the B-fragment accesses (row rho(g)) bank-conflict free (a double2 at rows 2tq, 2tq+1 collided 2-way).
                const int rho_g = g ^ (g >> 2);                                  // rho(g):  0 1 2 3 5 4 7 6
                const int rho0 = 2 * tq + (tq >> 1), rho1 = 2 * tq + 1 - (tq >> 1);   // rho(2tq) = 0 2 5 7, rho(2tq+1) = 1 3 4 6
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int rbase = 16 * warp + 8 * h;
                    double d[CT][2], bq[2 * MT];
#pragma unroll
                    for (int b = 0; b < CT; ++b) {
                        const double* xs = &T[(size_t)(8 * MT + 8 * b + g) * kPitch + rbase];
                        d[b][0] = xs[rho0];
                        d[b][1] = xs[rho1];
                    }
#pragma unroll
                    for (int ks = 0; ks < 2 * MT; ++ks) bq[ks] = T[(size_t)(4 * ks + tq) * kPitch + rbase + rho_g];
#pragma unroll
                    for (int ks = 0; ks < 2 * MT; ++ks)
#pragma unroll
                        for (int b = 0; b < CT; ++b) dmma(d[b][0], d[b][1], a1[ks][b], bq[ks]);
                    if (SOLVE && second) {       // Z = Y - Q*C2: the reference's second projection pass (project.m:35)
#pragma unroll
                        for (int ks = 0; ks < 2 * MT; ++ks)
#pragma unroll
                            for (int b = 0; b < CT; ++b) dmma(d[b][0], d[b][1], a2[ks][b], bq[ks]);
                    }
#pragma unroll
                    for (int b = 0; b < CT; ++b) {
                        double* xs = &T[(size_t)(8 * MT + 8 * b + g) * kPitch + rbase];
                        xs[rho0] = d[b][0];
                        xs[rho1] = d[b][1];
                        if (!SOLVE && p.Y && 8 * b + g < p.c) {
                            const long long r = r0 + rbase;
                            double* dst = p.Y + (long long)(8 * b + g) * p.ldY + r;
                            if (r + rho0 < p.n) dst[rho0] = d[b][0];
                            if (r + rho1 < p.n) dst[rho1] = d[b][1];
                        }
                    }
                }
                __syncwarp();                    // the warp's 16 rows of Y are in shared memory
                if (SOLVE) {
                    // ---- forward substitution along the row, q_j = (z_j - sum_{i<j} q_i R_ij) * (1/R_jj), same operation order as
                    //      k_trsolve.  Lane (row = lane & 15, half = lane >> 4) owns CW/2 columns of one of the warp's 16 rows; the
                    //      lane of the upper column half picks q_0..q_{HC-1} up from shared memory.
                    constexpr int HC = CW / 2;
                    const int row = 16 * warp + (lane & 15), half = lane >> 4;
                    const double* Rs = red;
                    const double* Rinv = Cs2 + (size_t)MT * 8 * CW;      // 1/R_jj, rounded once (as in k_trsolve)
                    double y[HC];
#pragma unroll
                    for (int j = 0; j < HC; ++j) y[j] = T[(size_t)(8 * MT + half * HC + j) * kPitch + row];
                    if (half == 0) {
#pragma unroll
                        for (int j = 0; j < HC; ++j) {
                            double sacc = y[j];
#pragma unroll
                            for (int i = 0; i < j; ++i) sacc = fma(-y[i], Rs[i * (CW + 1) + j], sacc);
                            y[j] = sacc * Rinv[j];
                            T[(size_t)(8 * MT + j) * kPitch + row] = y[j];
                        }
                    }
                    __syncwarp();
                    if (half == 1) {
                        double ql[HC];
#pragma unroll
                        for (int i = 0; i < HC; ++i) ql[i] = T[(size_t)(8 * MT + i) * kPitch + row];
#pragma unroll
                        for (int j = 0; j < HC; ++j) {
                            double sacc = y[j];
#pragma unroll
                            for (int i = 0; i < HC; ++i) sacc = fma(-ql[i], Rs[i * (CW + 1) + HC + j], sacc);
#pragma unroll
                            for (int i = 0; i < j; ++i) sacc = fma(-y[i], Rs[(HC + i) * (CW + 1) + HC + j], sacc);
                            y[j] = sacc * Rinv[HC + j];
                        }
                    }
                    const bool ok = r0 + row < p.n;
#pragma unroll
                    for (int j = 0; j < HC; ++j) {
                        const int col = half * HC + j;
                        if (ok && col < p.c) p.Y[(long long)col * p.ldY + r0 + row] = y[j];
                    }
                }
            }
            // ---- contraction over the 16 rows of this warp: S += [Q Y]' Y   (skipped when the caller wants no S: a pure update)
            if (p.partials)
#pragma unroll
            for (int u = 0; u < (SOLVE ? 0 : 4); ++u) {
                const int row = 16 * warp + 4 * u + tq;
                double bf[CT], af[NRT];
#pragma unroll
                for (int b = 0; b < CT; ++b) bf[b] = T[(size_t)(8 * MT + 8 * b + g) * kPitch + row];
#pragma unroll
                for (int a = 0; a < NRT; ++a) {
                    const int rt = RT0 + a;
                    af[a] = (rt >= MT) ? bf[rt - MT] : T[(size_t)(8 * rt + g) * kPitch + row];
                }
#pragma unroll
                for (int a = 0; a < NRT; ++a)
#pragma unroll
                    for (int b = 0; b < CT; ++b) dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
            }
            if (MODE != MODE_COEFF) fence_proxy_async();      // our generic-proxy writes vs the next TMA refill of this stage
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        // ---- CTA reduction over the consumer warps (fixed order)
        double* mine = red + (size_t)warp * NRT * CT * 64;
#pragma unroll
        for (int a = 0; a < (SOLVE ? 0 : NRT); ++a)
#pragma unroll
            for (int b = 0; b < CT; ++b) {
                double* tile = mine + (a * CT + b) * 64;
                tile[g * 8 + 2 * tq] = acc[a][b][0];
                tile[g * 8 + 2 * tq + 1] = acc[a][b][1];
            }
    }
    if (SOLVE || !p.partials) return;
    __syncthreads();
    constexpr int ELEMS = NRT * CT * 64;
    double* out = p.partials + (size_t)blockIdx.x * ELEMS;
    for (int e = tid; e < ELEMS; e += kTileThreads) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) sum += red[(size_t)w * ELEMS + e];
        out[e] = sum;
    }
}

template <int MT, int CT, int MODE>
__device__ __forceinline__ double tile_partial_sum(const double* __restrict__ partials, int nparts, int e, int lane) {
    constexpr int RT0 = (MODE == MODE_UPDATE_GRAM) ? MT : 0;
    constexpr int ELEMS = (MT + CT - RT0) * CT * 64;
    double sum = 0.0;
    for (int b0 = lane; b0 < nparts; b0 += 32 * 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int b = b0 + 32 * q;
            v[q] = b < nparts ? partials[(size_t)b * ELEMS + e] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) sum += v[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    return sum;
}

template <int MT, int CT, int MODE>
__device__ __forceinline__ void tile_finalize_ar(const double* __restrict__ partials, int nparts, double* __restrict__ S, int ldS, int M,
                                                 int c, const ArArgs& ar) {
    constexpr int RT0 = (MODE == MODE_UPDATE_GRAM) ? MT : 0;
    constexpr int NRT = MT + CT - RT0;
    constexpr int ELEMS = NRT * CT * 64;
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int buf = (int)(ar.seq % kMboxBufs);
    const size_t slot = ((size_t)buf * kMaxPeers + ar.me) * kMboxSlot;
    if (e < ELEMS) {
        const int tl = e >> 6, i = (e >> 3) & 7, j = e & 7;
        const int rt = RT0 + tl / CT, ct = tl % CT;
        const int col = 8 * ct + j;
        int rowS;
        bool valid;
        if (rt < MT) { rowS = 8 * rt + i; valid = rowS < M; }
        else { const int k = 8 * (rt - MT) + i; rowS = M + k; valid = k < c; }
        if (valid && col < c) {
            const double sum = tile_partial_sum<MT, CT, MODE>(partials, nparts, e, lane);
            if (lane < ar.P) ar.peers.mbox[lane][slot + (size_t)col * ldS + rowS] = sum;       // one peer per lane
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_last, s_ok;
    if (threadIdx.x == 0) {
        s_last = (atomicAdd(ar.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        s_ok = 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *ar.ticket = 0;
    __threadfence_system();
    if (threadIdx.x < ar.P) st_release_sys(ar.peers.flags[threadIdx.x] + kFlagAllreduce + buf * kMaxPeers + ar.me, ar.seq);
    if (threadIdx.x < ar.P && !spin_until(ar.my_flags + kFlagAllreduce + buf * kMaxPeers + threadIdx.x, ar.seq, ar.err)) s_ok = 0;
    __syncthreads();
    if (!s_ok) return;
    const int count = ldS * c;
    for (int idx = threadIdx.x; idx < count; idx += blockDim.x) {
        const int rowS = idx % ldS;
        if (MODE == MODE_UPDATE_GRAM && rowS < M) continue;                                  // only the block part is produced
        double s = 0.0;
        for (int r = 0; r < ar.P; ++r) s += __ldcv(ar.my_mbox + ((size_t)buf * kMaxPeers + r) * kMboxSlot + idx);
        S[idx] = s;
    }
}

// Deterministic second stage of the reduction: one warp per element of S sums the per-CTA partials in a fixed order
// (lanes stride over the CTAs with all loads in flight at once, fixed shuffle tree).  A separate small launch on purpose:
// done by the last CTA of k_tile it cost ~30 us of serialised L2 latency per pass.
// With a communicator (ar.P > 1) the same launch is the all-reduce as well (peer-memory mailbox, see p2p.cu): every element goes
// straight into slot [seq % 4][me] of EVERY rank's mailbox instead of into S; the last CTA to finish raises this rank's flag at
// all peers, waits for theirs and writes S = sum over ranks (rank order: identical bits everywhere).  One launch instead of
// finalize + all-reduce, and the local sums never make the round trip through S.
template <int MT, int CT, int MODE>
__global__ void __launch_bounds__(256)
k_tile_finalize(const double* __restrict__ partials, int nparts, double* __restrict__ S, int ldS, int M, int c,
                const int* __restrict__ pred, int want, const __grid_constant__ ArArgs ar) {
    if (pred && *pred != want) return;
    if (ar.P > 1) {
        tile_finalize_ar<MT, CT, MODE>(partials, nparts, S, ldS, M, c, ar);
        return;
    }
    constexpr int RT0 = (MODE == MODE_UPDATE_GRAM) ? MT : 0;
    constexpr int NRT = MT + CT - RT0;
    constexpr int ELEMS = NRT * CT * 64;
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= ELEMS) return;
    const int tl = e >> 6, i = (e >> 3) & 7, j = e & 7;
    const int rt = RT0 + tl / CT, ct = tl % CT;
    const int col = 8 * ct + j;
    int rowS;                                          // row of S: Q-part first, block-part after it
    bool valid;
    if (rt < MT) { rowS = 8 * rt + i; valid = rowS < M; }
    else { const int k = 8 * (rt - MT) + i; rowS = M + k; valid = k < c; }
    if (!valid || col >= c) return;
    const double sum = tile_partial_sum<MT, CT, MODE>(partials, nparts, e, lane);
    if (lane == 0) S[(size_t)col * ldS + rowS] = sum;
}

template <int MT, int CT, int MODE>
int launch_tile(calz_ctx* ctx, const TileArgs& a0, bool fuse_allreduce = false) {
    TileArgs a = a0;
    constexpr int SLOTS = 8 * (MT + CT);
    constexpr bool SOLVE = (MODE == MODE_UPDATE_SOLVE);
    constexpr int NRT = (MODE == MODE_UPDATE_GRAM || SOLVE) ? CT : MT + CT;
    const size_t stage_bytes = (size_t)SLOTS * kPitch * sizeof(double);
    const size_t red_doubles = SOLVE ? (size_t)(8 * CT) * (8 * CT + 1) + (size_t)MT * 8 * 8 * CT + 8 * CT : (size_t)kConsumerWarps * NRT * CT * 64;
    const size_t fixed = ((size_t)MT * 8 * 8 * CT + red_doubles) * sizeof(double) + 2 * 8 * sizeof(uint64_t) + 64;
    int dev_max = 0;
    cudaDeviceGetAttribute(&dev_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
    // aim for 2 CTAs per SM: 228 KB per SM minus 1 KB the driver reserves per resident CTA
    int sm_max = 0;
    cudaDeviceGetAttribute(&sm_max, cudaDevAttrMaxSharedMemoryPerMultiprocessor, ctx->device);
    const size_t budget = ((size_t)sm_max - 2 * 1024) / 2 - 512;
    int stages = 4;
    while (stages > 2 && fixed + stages * stage_bytes > budget) --stages;
    if (fixed + stages * stage_bytes > budget) {         // wide panels: one CTA per SM with as deep a ring as fits
        stages = 4;
        while (stages > 2 && fixed + stages * stage_bytes > (size_t)dev_max) --stages;
    }
    const size_t smem = fixed + stages * stage_bytes;
    if (smem > (size_t)dev_max) return set_error(ctx, CALZ_ERR_UNSUPPORTED, "tile kernel needs %zu B of shared memory", smem);
    auto kern = k_tile<MT, CT, MODE>;
    CALZ_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    per_sm = std::min(per_sm, 2);
    const long long ntiles = (a.n + kTileRows - 1) / kTileRows;
    const int grid = (int)std::max<long long>(1, std::min<long long>((long long)ctx->num_sms * per_sm, ntiles));
    const bool want_S = !SOLVE && a.S != nullptr;
    a.partials = nullptr;
    if (want_S) {
        CALZ_TRY(reserve(ctx, ctx->partials, (size_t)grid * NRT * CT * 64 * sizeof(double)));
        a.partials = (double*)ctx->partials.p;
    }
    a.ticket = ctx->ticket;
    kern<<<grid, kTileThreads, smem, ctx->stream>>>(a, stages);
    CALZ_LAUNCH_CHECK(ctx);
    if (!want_S) return CALZ_OK;
    ArArgs ar{};
    ar.P = 1;
    if (fuse_allreduce) p2p_next_allreduce(ctx, &ar);
    k_tile_finalize<MT, CT, MODE><<<(NRT * CT * 64 + 7) / 8, 256, 0, ctx->stream>>>(a.partials, grid, a.S, a.ldS, a.M, a.c, a.pred, a.want, ar);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

bool aligned16(const void* p, long long ld) { return ((uintptr_t)p % 16 == 0) && (ld % 2 == 0); }

}  // namespace

// The fused passes apply to any row count and any alignment (a rank whose blocks are not 16-byte aligned, or shorter than one
// tile, stages its tiles with plain loads instead of TMA bulk copies): which kernels run -- and therefore which collectives are
// issued -- depends on the block SHAPE only, never on rank-local properties.
bool tile_path_ok(int64_t n, const double* Q, int64_t ldQ, int M, const double* X, int64_t ldX, int c, const double* Y, int64_t ldY) {
    (void)Q; (void)ldQ; (void)X; (void)ldX; (void)Y; (void)ldY;
    return n >= 1 && M >= 1 && M <= 16 && c >= 1 && c <= 16;
}
// widest panel of Q one tile pass takes against a block of c columns (shared memory: 8*(MT+CT) column slots per stage)
int tile_panel_width(int c) { return c <= 8 ? 48 : (c <= 16 ? 32 : 0); }
static int tile_use_tma(const TileArgs& a) { return aligned16(a.Q, a.ldQ) && aligned16(a.X, a.ldX) ? 1 : 0; }

// mode: 0 = COEFF (S = [Q X]'X), 1 = UPDATE_FULL (Y = X - Q*C, S = [Q Y]'Y), 2 = UPDATE_GRAM (Y = X - Q*C, S rows [M,M+c) = Y'Y)
int tile_pass(calz_ctx* ctx, int mode, int64_t n, const double* Q, int64_t ldQ, int M, const double* X, int64_t ldX, int c,
              const double* C_dev, int ldC, double* Y, int64_t ldY, double* S_dev, int ldS, const int* pred, int want, bool allreduce) {
    TileArgs a{};
    a.n = n; a.Q = Q; a.ldQ = ldQ; a.M = M; a.X = X; a.ldX = ldX; a.c = c; a.Y = Y; a.ldY = ldY; a.C = C_dev; a.ldC = ldC;
    a.S = S_dev; a.ldS = ldS; a.pred = pred; a.want = want;
    a.use_tma = tile_use_tma(a);
    const int MT = (M + 7) / 8, CT = (c + 7) / 8;
    if (M < 1 || c < 1 || CT > 2 || M > tile_panel_width(c) || (mode == 1 && MT > 2))
        return set_error(ctx, CALZ_ERR_BADARG, "tile_pass: panel of %d columns against a block of %d", M, c);
    int st;
    const bool want_S = S_dev != nullptr;
    // with a communicator and the peer-memory mailbox available the finalize launch is the all-reduce too.  Not for a predicated
    // pass: a skipped launch would leave the peers waiting (the generic all-reduce below always runs).
    const bool fuse_ar = want_S && allreduce && ctx->nranks > 1 && !pred && ldS == M + c && p2p_allreduce_ok(ctx, (size_t)ldS * c) &&
                         ctx->opt_fused_allreduce;
#define CALZ_TILE(MTv, CTv)                                                         \
    st = mode == 0 ? launch_tile<MTv, CTv, MODE_COEFF>(ctx, a, fuse_ar)             \
       : mode == 1 ? launch_tile<MTv, CTv, MODE_UPDATE_FULL>(ctx, a, fuse_ar)       \
                   : launch_tile<MTv, CTv, MODE_UPDATE_GRAM>(ctx, a, fuse_ar)
    // wide panels (multi-block projections, 'full' re-orthogonalisation): coefficient and update+Gram passes only
#define CALZ_TILE_WIDE(MTv, CTv)                                                    \
    st = mode == 0 ? launch_tile<MTv, CTv, MODE_COEFF>(ctx, a, fuse_ar)             \
                   : launch_tile<MTv, CTv, MODE_UPDATE_GRAM>(ctx, a, fuse_ar)
    if (MT == 1 && CT == 1) { CALZ_TILE(1, 1); }
    else if (MT == 2 && CT == 1) { CALZ_TILE(2, 1); }
    else if (MT == 1 && CT == 2) { CALZ_TILE(1, 2); }
    else if (MT == 2 && CT == 2) { CALZ_TILE(2, 2); }
    else if (CT == 1) {
        if (MT == 3) { CALZ_TILE_WIDE(3, 1); } else if (MT == 4) { CALZ_TILE_WIDE(4, 1); } else { CALZ_TILE_WIDE(6, 1); }
    } else {
        if (MT == 3) { CALZ_TILE_WIDE(3, 2); } else { CALZ_TILE_WIDE(4, 2); }
    }
#undef CALZ_TILE_WIDE
#undef CALZ_TILE
    CALZ_TRY(st);
    if (want_S && allreduce && ctx->nranks > 1 && !fuse_ar) {
        // a predicated-off pass leaves S untouched on every rank alike, so the collective stays consistent
        if (mode == 2) {
            if (ldS != M + c) return set_error(ctx, CALZ_ERR_BADARG, "tile_pass: dense S expected");
            // only the block-part rows are produced: reduce them column by column would cost c collectives; reduce all
        }
        CALZ_TRY(allreduce_sum(ctx, S_dev, (size_t)ldS * c));
    }
    return CALZ_OK;
}

// QZ = ((X - Q*C1) - [*flag2] Q*C2) / R in ONE pass over [Q | X].  Same per-row arithmetic as the UPDATE passes followed by
// k_trsolve.
int tile_update_solve(calz_ctx* ctx, int64_t n, const double* Q, int64_t ldQ, int M, const double* X, int64_t ldX, int c,
                      const double* C1_dev, int ldC1, const double* C2_dev, int ldC2, const int* flag2, const double* R_dev, double* QZ,
                      int64_t ldQZ) {
    TileArgs a{};
    a.n = n; a.Q = Q; a.ldQ = ldQ; a.M = M; a.X = X; a.ldX = ldX; a.c = c; a.Y = QZ; a.ldY = ldQZ; a.C = C1_dev; a.ldC = ldC1;
    a.C2 = C2_dev; a.ldC2 = ldC2; a.flag2 = flag2; a.R = R_dev;
    a.use_tma = tile_use_tma(a);
    const int MT = (M + 7) / 8, CT = (c + 7) / 8;
    if (MT == 1 && CT == 1) return launch_tile<1, 1, MODE_UPDATE_SOLVE>(ctx, a);
    if (MT == 2 && CT == 1) return launch_tile<2, 1, MODE_UPDATE_SOLVE>(ctx, a);
    if (MT == 1 && CT == 2) return launch_tile<1, 2, MODE_UPDATE_SOLVE>(ctx, a);
    return launch_tile<2, 2, MODE_UPDATE_SOLVE>(ctx, a);
}

}  // namespace calz
