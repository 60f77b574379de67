// Peer-memory collectives over NVLink/NVSwitch (CUDA IPC), one process per GPU.
//
// The two exchanges of an outer step are tiny or nearest-neighbour, i.e. latency-bound: measured on 8xB200, NCCL takes
// 26 us for the all-reduce of a 17x8 coefficient block and 80 us for the 4 MB halo exchange.  Here both are plain
// stores into the peers' memory followed by a flag (release at system scope), and a bounded spin on the local flags:
//   * all-reduce ("mailbox"): every rank stores its contribution into slot [seq%4][me] of EVERY peer's mailbox, raises
//     flag [seq%4][me] there, waits for the P flags of its own mailbox and sums the P slots in rank order -- the same
//     order on every rank, hence bit-identical results everywhere and run-to-run (NCCL guarantees neither);
//   * halo: the owner pushes the level-s boundary rows of the new start vector straight into the ghost zone of the
//     neighbour's basis workspace (no pack buffer, no receive copy) and raises a flag; the consumer acknowledges after its
//     first SpMV so that the next push cannot overwrite ghosts that are still being read.
// All spins are bounded; on timeout a device error flag is raised and every later kernel of the library bails out.
// NCCL stays as the bootstrap (handle exchange) and as the fallback when IPC is unavailable or a message is too large.
#include <string.h>

#include "matrix.h"
#include "p2p_dev.cuh"

namespace calz {

namespace {

// in place: data[i] = sum_r contribution_r[i], r in rank order
__global__ void __launch_bounds__(256)
k_allreduce_p2p(double* data, int count, int P, int me, unsigned long long seq, PeerPtrs peers, double* my_mbox,
                unsigned long long* my_flags, int* err) {
    const int buf = (int)(seq % kMboxBufs);
    const size_t slot = ((size_t)buf * kMaxPeers + me) * kMboxSlot;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const double v = data[i];
        for (int q = 0; q < P; ++q) peers.mbox[q][slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P) st_release_sys(peers.flags[threadIdx.x] + kFlagAllreduce + buf * kMaxPeers + me, seq);
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < P && !spin_until(my_flags + kFlagAllreduce + buf * kMaxPeers + threadIdx.x, seq, err)) ok = 0;
    __syncthreads();
    if (!ok) return;
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < P; ++r) s += __ldcv(my_mbox + ((size_t)buf * kMaxPeers + r) * kMboxSlot + i);
        data[i] = s;
    }
}

struct HaloPush {
    int npeer;
    int peer[kMaxPeers];
    double* dst[kMaxPeers];                 // where my rows start in the peer's workspace column 0
    const int* idx[kMaxPeers];              // local indices to gather (NULL: contiguous from src_off)
    long long src_off[kMaxPeers];
    long long count[kMaxPeers];
    unsigned long long ack_wait[kMaxPeers]; // sequence number of my previous push to that peer: must be consumed first
    unsigned long long* peer_flags[kMaxPeers];
};

// grid (chunks, npeer): copy my boundary rows into the peer's ghost zone; the last CTA per peer raises the data flag
__global__ void __launch_bounds__(256)
k_halo_push(const double* __restrict__ w, HaloPush h, int me, unsigned long long seq, unsigned long long* my_flags,
            unsigned int* tickets, int* err) {
    const int k = blockIdx.y, q = h.peer[k];
    // the peer must have consumed the ghosts of the previous exchange (its ack lands in my flag array)
    __shared__ int ok;
    if (threadIdx.x == 0) ok = spin_until(my_flags + kFlagHaloAck + q, h.ack_wait[k], err) ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    const long long n = h.count[k];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = h.idx[k] ? w[h.idx[k][i]] : w[h.src_off[k] + i];
        h.dst[k][i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(tickets + k, 1u) == gridDim.x - 1) {
            tickets[k] = 0;
            __threadfence_system();
            st_release_sys(h.peer_flags[k] + kFlagHaloData + me, seq);
        }
    }
}

struct HaloWait {
    int n;
    int peer[kMaxPeers];
    unsigned long long* peer_flags[kMaxPeers];
};

__global__ void k_halo_wait(HaloWait hw, unsigned long long seq, unsigned long long* my_flags, int* err) {
    if (threadIdx.x < hw.n) spin_until(my_flags + kFlagHaloData + hw.peer[threadIdx.x], seq, err);
}

__global__ void k_halo_ack(HaloWait hw, int me, unsigned long long seq) {
    if (threadIdx.x < hw.n) st_release_sys(hw.peer_flags[threadIdx.x] + kFlagHaloAck + me, seq);
}

// exchange `bytes` per rank with every peer through NCCL send/recv (bit-exact: no arithmetic)
int exchange_blobs(calz_ctx* ctx, const void* mine, size_t bytes, std::vector<std::vector<unsigned char>>& all) {
    const int P = ctx->nranks;
    const size_t nd = (bytes + 7) / 8;
    double *d_send = nullptr, *d_recv = nullptr;
    CALZ_CUDA(ctx, cudaMalloc(&d_send, nd * sizeof(double)));
    CALZ_CUDA(ctx, cudaMalloc(&d_recv, nd * P * sizeof(double)));
    std::vector<unsigned char> tmp(nd * 8, 0);
    memcpy(tmp.data(), mine, bytes);
    CALZ_CUDA(ctx, cudaMemcpy(d_send, tmp.data(), nd * 8, cudaMemcpyHostToDevice));
    CALZ_NCCL(ctx, ctx->nccl->GroupStart());
    for (int q = 0; q < P; ++q) {
        if (q == ctx->rank) continue;
        CALZ_NCCL(ctx, ctx->nccl->Send(d_send, nd, ncclFloat64, q, ctx->comm, ctx->stream));
        CALZ_NCCL(ctx, ctx->nccl->Recv(d_recv + (size_t)q * nd, nd, ncclFloat64, q, ctx->comm, ctx->stream));
    }
    CALZ_NCCL(ctx, ctx->nccl->GroupEnd());
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<unsigned char> host(nd * 8 * P);
    CALZ_CUDA(ctx, cudaMemcpy(host.data(), d_recv, host.size(), cudaMemcpyDeviceToHost));
    cudaFree(d_send);
    cudaFree(d_recv);
    all.assign(P, std::vector<unsigned char>(bytes));
    for (int q = 0; q < P; ++q)
        memcpy(all[q].data(), q == ctx->rank ? (const unsigned char*)mine : host.data() + (size_t)q * nd * 8, bytes);
    return CALZ_OK;
}

}  // namespace

// ---- set-up of the per-context mailbox (called from calz_comm_init after NCCL is up); failure => NCCL-only operation
int p2p_setup(calz_ctx* ctx) {
    P2P& p = ctx->p2p;
    p.enabled = false;
    const int P = ctx->nranks;
    if (P < 2 || P > kMaxPeers || !ctx->opt_p2p) return CALZ_OK;
    const size_t mbox_bytes = (size_t)kMboxBufs * kMaxPeers * kMboxSlot * sizeof(double);
    const size_t flag_bytes = (size_t)kFlagCount * sizeof(unsigned long long);
    CALZ_CUDA(ctx, cudaMalloc(&p.base, mbox_bytes + flag_bytes + 256));
    CALZ_CUDA(ctx, cudaMemset(p.base, 0, mbox_bytes + flag_bytes + 256));
    p.mbox = (double*)p.base;
    p.flags = (unsigned long long*)((char*)p.base + mbox_bytes);
    CALZ_CUDA(ctx, cudaMalloc(&p.err, 64));
    CALZ_CUDA(ctx, cudaMemset(p.err, 0, 64));
    CALZ_CUDA(ctx, cudaMalloc(&p.tickets, 2 * kMaxPeers * sizeof(unsigned int)));
    CALZ_CUDA(ctx, cudaMemset(p.tickets, 0, 2 * kMaxPeers * sizeof(unsigned int)));
    cudaIpcMemHandle_t mine;
    cudaError_t e = cudaIpcGetMemHandle(&mine, p.base);
    int ok = (e == cudaSuccess) ? 1 : 0;
    if (!ok) cudaGetLastError();
    std::vector<std::vector<unsigned char>> all;
    struct Blob { cudaIpcMemHandle_t h; int ok; } blob;
    memset(&blob, 0, sizeof(blob));
    blob.h = mine;
    blob.ok = ok;
    CALZ_TRY(exchange_blobs(ctx, &blob, sizeof(blob), all));
    bool all_ok = true;
    for (int q = 0; q < P; ++q) all_ok &= ((Blob*)all[q].data())->ok != 0;
    for (int q = 0; q < P && all_ok; ++q) {
        if (q == ctx->rank) {
            p.peer_base[q] = p.base;
        } else {
            void* ptr = nullptr;
            e = cudaIpcOpenMemHandle(&ptr, ((Blob*)all[q].data())->h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                all_ok = false;
                break;
            }
            p.peer_base[q] = ptr;
        }
    }
    // every rank must agree (a one-sided fallback would dead-lock the collectives): reduce the verdict over NCCL
    double verdict = all_ok ? 0.0 : 1.0, *d_v = nullptr;
    CALZ_CUDA(ctx, cudaMalloc(&d_v, sizeof(double)));
    CALZ_CUDA(ctx, cudaMemcpy(d_v, &verdict, sizeof(double), cudaMemcpyHostToDevice));
    CALZ_NCCL(ctx, ctx->nccl->AllReduce(d_v, d_v, 1, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CALZ_CUDA(ctx, cudaMemcpy(&verdict, d_v, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_v);
    p.enabled = verdict == 0.0;
    p.seq_allreduce = 0;
    return CALZ_OK;
}

void p2p_teardown(calz_ctx* ctx) {
    P2P& p = ctx->p2p;
    for (int q = 0; q < ctx->nranks && q < kMaxPeers; ++q)
        if (p.peer_base[q] && q != ctx->rank) cudaIpcCloseMemHandle(p.peer_base[q]);
    if (p.base) cudaFree(p.base);
    if (p.err) cudaFree(p.err);
    if (p.tickets) cudaFree(p.tickets);
    p = P2P();
}

static size_t mbox_bytes_total() { return (size_t)kMboxBufs * kMaxPeers * kMboxSlot * sizeof(double); }

bool p2p_allreduce_ok(const calz_ctx* ctx, size_t count) { return ctx->p2p.enabled && count <= (size_t)kMboxSlot; }

int p2p_allreduce(calz_ctx* ctx, double* dev, size_t count) {
    P2P& p = ctx->p2p;
    PeerPtrs pp{};
    for (int q = 0; q < ctx->nranks; ++q) {
        pp.mbox[q] = (double*)p.peer_base[q];
        pp.flags[q] = (unsigned long long*)((char*)p.peer_base[q] + mbox_bytes_total());
    }
    const unsigned long long seq = ++p.seq_allreduce;
    k_allreduce_p2p<<<1, 256, 0, ctx->stream>>>(dev, (int)count, ctx->nranks, ctx->rank, seq, pp, p.mbox, p.flags, p.err);
    CALZ_LAUNCH_CHECK(ctx);
    return CALZ_OK;
}

void p2p_next_allreduce(calz_ctx* ctx, ArArgs* a) {
    P2P& p = ctx->p2p;
    a->P = ctx->nranks;
    a->me = ctx->rank;
    for (int q = 0; q < ctx->nranks; ++q) {
        a->peers.mbox[q] = (double*)p.peer_base[q];
        a->peers.flags[q] = (unsigned long long*)((char*)p.peer_base[q] + mbox_bytes_total());
    }
    a->seq = ++p.seq_allreduce;
    a->my_mbox = p.mbox;
    a->my_flags = p.flags;
    a->err = p.err;
    a->ticket = p.tickets + kMaxPeers;          // its own ticket word (the halo push uses [0, kMaxPeers))
}

int p2p_check(calz_ctx* ctx) {      // after a synchronisation: did any bounded spin give up?
    if (!ctx->p2p.enabled) return CALZ_OK;
    int e = 0;
    CALZ_CUDA(ctx, cudaMemcpy(&e, ctx->p2p.err, sizeof(int), cudaMemcpyDeviceToHost));
    if (e) return set_error(ctx, CALZ_ERR_NCCL, "peer-memory collective timed out waiting for another rank");
    return CALZ_OK;
}

// ---- per-matrix halo plan: open the peers' basis workspaces, learn where my rows land there
int p2p_halo_setup(calz_mat* m) {
    calz_ctx* ctx = m->ctx;
    m->p2p_halo = false;
    if (!ctx->p2p.enabled) return CALZ_OK;
    const int P = ctx->nranks;
    struct Blob { cudaIpcMemHandle_t h; long long recv_off[kMaxPeers]; long long ldW; int ok; int w_pad; } blob;
    memset(&blob, 0, sizeof(blob));
    cudaError_t e = cudaIpcGetMemHandle(&blob.h, m->d_W_alloc);
    blob.w_pad = m->W_pad;
    blob.ldW = m->ldW;
    blob.ok = (e == cudaSuccess);
    if (!blob.ok) cudaGetLastError();
    for (int q = 0; q < P; ++q) blob.recv_off[q] = m->recv_off[q];
    std::vector<std::vector<unsigned char>> all;
    CALZ_TRY(exchange_blobs(ctx, &blob, sizeof(blob), all));
    bool all_ok = true;
    for (int q = 0; q < P; ++q) all_ok &= ((Blob*)all[q].data())->ok != 0;
    m->peer_W.assign(P, nullptr);
    m->peer_W_base.assign(P, nullptr);
    m->peer_dst_off.assign(P, 0);
    m->peer_ldW.assign(P, 0);
    for (int q = 0; q < P && all_ok; ++q) {
        if (q == ctx->rank || (m->send_cnt[q] == 0 && m->recv_cnt[q] == 0)) continue;
        void* ptr = nullptr;
        e = cudaIpcOpenMemHandle(&ptr, ((Blob*)all[q].data())->h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); all_ok = false; break; }
        m->peer_W[q] = (double*)ptr + ((Blob*)all[q].data())->w_pad;      // the peer's d_W (its allocation + alignment shift)
        m->peer_W_base[q] = ptr;
        m->peer_dst_off[q] = ((Blob*)all[q].data())->recv_off[ctx->rank];      // where q receives MY rows
        m->peer_ldW[q] = ((Blob*)all[q].data())->ldW;
    }
    double verdict = all_ok ? 0.0 : 1.0, *d_v = nullptr;
    CALZ_CUDA(ctx, cudaMalloc(&d_v, sizeof(double)));
    CALZ_CUDA(ctx, cudaMemcpy(d_v, &verdict, sizeof(double), cudaMemcpyHostToDevice));
    CALZ_NCCL(ctx, ctx->nccl->AllReduce(d_v, d_v, 1, ncclFloat64, ncclSum, ctx->comm, ctx->stream));
    CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    CALZ_CUDA(ctx, cudaMemcpy(&verdict, d_v, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_v);
    m->p2p_halo = verdict == 0.0;
    // sequence numbers of the halo flags are per context (flags live in the context mailbox): start above whatever
    // an earlier matrix of this context used, identically on all ranks
    return CALZ_OK;
}

void p2p_halo_teardown(calz_mat* m) {
    for (size_t q = 0; q < m->peer_W_base.size(); ++q)
        if (m->peer_W_base[q]) cudaIpcCloseMemHandle(m->peer_W_base[q]);
    m->peer_W.clear();
    m->peer_W_base.clear();
}

// push my boundary rows of workspace column `col` into the peers' ghost zones (same column), wait for theirs
int p2p_halo_exchange(calz_mat* m, double* W, int col) {
    const double* w = W + (long long)col * m->ldW;
    calz_ctx* ctx = m->ctx;
    P2P& p = ctx->p2p;
    const int P = ctx->nranks;
    const unsigned long long seq = ++p.seq_halo;
    HaloPush h{};
    HaloWait hw{};
    long long maxn = 0;
    for (int q = 0; q < P; ++q) {
        unsigned long long* pf = (unsigned long long*)((char*)p.peer_base[q] + mbox_bytes_total());
        if (q != ctx->rank && m->send_cnt[q]) {
            const int k = h.npeer++;
            h.peer[k] = q;
            h.dst[k] = m->peer_W[q] + (long long)col * m->peer_ldW[q] + m->peer_dst_off[q];
            h.idx[k] = m->send_contig[q] ? nullptr : m->d_send_idx + m->send_off[q];
            h.src_off[k] = m->own_off + (m->send_glob[q][0] - m->row_lo);
            h.count[k] = m->send_cnt[q];
            h.ack_wait[k] = p.last_push[q];
            p.last_push[q] = seq;
            h.peer_flags[k] = pf;
            maxn = std::max<long long>(maxn, m->send_cnt[q]);
        }
        if (q != ctx->rank && m->recv_cnt[q]) {
            const int k = hw.n++;
            hw.peer[k] = q;
            hw.peer_flags[k] = pf;
        }
    }
    if (h.npeer) {
        const int chunks = (int)std::max<long long>(1, std::min<long long>(64, (maxn + 4095) / 4096));
        k_halo_push<<<dim3(chunks, h.npeer), 256, 0, ctx->stream>>>(w, h, ctx->rank, seq, p.flags, p.tickets, p.err);
        CALZ_LAUNCH_CHECK(ctx);
    }
    if (hw.n) {
        k_halo_wait<<<1, 32, 0, ctx->stream>>>(hw, seq, p.flags, p.err);
        CALZ_LAUNCH_CHECK(ctx);
    }
    return CALZ_OK;
}

// after the first SpMV of the block: the ghosts of column 0 are no longer needed, tell the owners
int p2p_halo_ack(calz_mat* m) {
    calz_ctx* ctx = m->ctx;
    P2P& p = ctx->p2p;
    HaloWait hw{};
    for (int q = 0; q < ctx->nranks; ++q)
        if (q != ctx->rank && m->recv_cnt[q]) {
            const int k = hw.n++;
            hw.peer[k] = q;
            hw.peer_flags[k] = (unsigned long long*)((char*)p.peer_base[q] + mbox_bytes_total());
        }
    if (hw.n) {
        k_halo_ack<<<1, 32, 0, ctx->stream>>>(hw, ctx->rank, p.seq_halo);
        CALZ_LAUNCH_CHECK(ctx);
    }
    return CALZ_OK;
}

}  // namespace calz
