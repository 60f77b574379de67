// Host-side ingest of the sparse matrix: row partition, level-s ghost closure (PA1), local renumbering,
// CSR / SELL-32-sigma construction, exchange plan, device upload.
//
// The reference has no counterpart (single MATLAB process, built-in sparse type): this is the "other data
// structure" SpMV.m:3-5 leaves room for, and the partition scheme is the one of Hoemmen's thesis that
// ca_lanczos.m:3-5 cites.
#include <string.h>

#include <algorithm>
#include <map>
#include <numeric>

#include "matrix.h"

using namespace calz;

namespace {

// BFS level sets of the pattern graph starting from the owned rows.  level[j] = smallest k <= s with
// j in R_k, else -1.  Rows outside [row_begin,row_end) that would have to be expanded -> CALZ_ERR_CLOSURE.
int level_sets_host(calz_ctx* ctx, int64_t n_glob, int64_t row_begin, int64_t row_end, const int64_t* rowptr,
                    const int32_t* colind, int64_t lo, int64_t hi, int s, std::vector<int8_t>& level,
                    std::vector<std::vector<int64_t>>* by_level, int* covered = nullptr) {
    // covered != NULL: instead of failing when a frontier row is not supplied, report the deepest level whose sets are
    // complete (the caller then settles for a shallower ghost closure); levels beyond it are reset to -1.
    if (covered) *covered = s;
    if (s > 120) return set_error(ctx, CALZ_ERR_BADARG, "s=%d too large", s);
    level.assign((size_t)n_glob, (int8_t)-1);
    for (int64_t i = lo; i < hi; ++i) level[i] = 0;
    std::vector<int64_t> frontier, next;
    if (by_level) by_level->assign(s + 1, {});
    bool first = true;
    for (int k = 1; k <= s; ++k) {
        next.clear();
        auto expand = [&](int64_t i) -> bool {
            if (i < row_begin || i >= row_end) return false;
            for (int64_t e = rowptr[i - row_begin]; e < rowptr[i - row_begin + 1]; ++e) {
                int64_t j = colind[e];
                if (level[j] < 0) {
                    level[j] = (int8_t)k;
                    next.push_back(j);
                }
            }
            return true;
        };
        if (first) {
            for (int64_t i = lo; i < hi; ++i)
                if (!expand(i)) return set_error(ctx, CALZ_ERR_CLOSURE, "owned row %lld not supplied", (long long)i);
            first = false;
        } else {
            bool short_of_rows = false;
            for (int64_t i : frontier)
                if (!expand(i)) {
                    if (!covered)
                        return set_error(ctx, CALZ_ERR_CLOSURE,
                                         "row %lld (ghost level %d) is needed but rows [%lld,%lld) were supplied",
                                         (long long)i, k - 1, (long long)row_begin, (long long)row_end);
                    short_of_rows = true;
                    break;
                }
            if (short_of_rows) {                      // level k is incomplete: forget it
                for (int64_t j : next) level[j] = (int8_t)-1;
                *covered = k - 1;
                break;
            }
        }
        std::sort(next.begin(), next.end());
        if (by_level) (*by_level)[k] = next;
        frontier.swap(next);
        if (frontier.empty()) break;
    }
    return CALZ_OK;
}

template <class T>
int upload(calz_ctx* ctx, T** dptr, const std::vector<T>& h) {
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CALZ_CUDA(ctx, cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) CALZ_CUDA(ctx, cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return CALZ_OK;
}

}  // namespace

extern "C" {

int calz_partition_bounds(int64_t n, int P, int64_t* bounds) {
    if (n < 0 || P < 1 || !bounds) return set_error(nullptr, CALZ_ERR_BADARG, "calz_partition_bounds: bad arguments");
    for (int p = 0; p <= P; ++p) bounds[p] = (int64_t)(((__int128)p * n) / P);
    return CALZ_OK;
}

int calz_level_sets(int64_t n_glob, int64_t row_begin, int64_t row_end, const int64_t* rowptr,
                    const int32_t* colind, int64_t lo, int64_t hi, int s, int32_t* level_out) {
    if (!rowptr || !colind || !level_out || lo < 0 || hi > n_glob || lo > hi || s < 0)
        return set_error(nullptr, CALZ_ERR_BADARG, "calz_level_sets: bad arguments");
    std::vector<int8_t> level;
    CALZ_TRY(level_sets_host(nullptr, n_glob, row_begin, row_end, rowptr, colind, lo, hi, s, level, nullptr));
    for (int64_t i = 0; i < n_glob; ++i) level_out[i] = level[i];
    return CALZ_OK;
}

int calz_mat_destroy(calz_mat* m) {
    if (!m) return CALZ_OK;
    if (m->ctx) cudaStreamSynchronize(m->ctx->stream);
    p2p_halo_teardown(m);
    void* ptrs[] = {m->d_slice_pat, m->d_gridbar, m->d_long_row, m->d_long_seg0, m->d_long_segptr, m->d_long_segrow, m->d_long_col, m->d_long_val, m->d_long_part,
                    m->d_xs_off, m->d_codes, m->d_dict, m->d_send_idx, m->d_send_buf, m->d_rowptr, m->d_colind, m->d_val, m->d_slice_ptr,
                    m->d_sell_col, m->d_sell_val, m->d_perm, m->d_W_alloc};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete m;
    return CALZ_OK;
}

int calz_mat_create_csr(calz_ctx* ctx, int64_t n_glob, int64_t row_begin, int64_t row_end, const int64_t* rowptr,
                        const int32_t* colind, const double* val, int s_max, int layout, calz_mat** out) {
    if (!ctx || !out || !rowptr || !colind || !val || n_glob <= 0 || s_max < 1 || row_begin < 0 ||
        row_end > n_glob || row_begin > row_end)
        return set_error(ctx, CALZ_ERR_BADARG, "calz_mat_create_csr: bad arguments");
    *out = nullptr;
    CALZ_CUDA(ctx, cudaSetDevice(ctx->device));
    struct SetupGuard { calz_ctx* c; explicit SetupGuard(calz_ctx* c_) : c(c_) { c->setup_depth++; } ~SetupGuard() { c->setup_depth--; } } setup_guard(ctx);
    const int P = ctx->nranks, me = ctx->rank;
    std::vector<int64_t> bounds(P + 1);
    calz_partition_bounds(n_glob, P, bounds.data());
    calz_mat* m = new calz_mat();
    m->ctx = ctx;
    m->n_glob = n_glob;
    m->row_lo = bounds[me];
    m->row_hi = bounds[me + 1];
    m->n_own = m->row_hi - m->row_lo;
    m->s_max = s_max;
    const int64_t lo = m->row_lo, hi = m->row_hi;

    // ---- local index space: ascending global index over R_s(p)
    std::vector<int8_t> level;            // per GLOBAL row (P>1 only)
    std::vector<int64_t> loc2glob;        // ghosts below | owned | ghosts above  (only ghosts stored: ghost_glob)
    std::vector<int8_t> loc_level;
    if (P == 1) {
        if (row_begin != 0 || row_end != n_glob) {
            delete m;
            return set_error(ctx, CALZ_ERR_CLOSURE, "single rank: all rows [0,n) must be supplied");
        }
        m->n_loc = n_glob;
        m->own_off = 0;
        m->halo_level = s_max;
    } else {
        // ---- depth L of the ghost closure (= MPK steps per halo exchange).  Explicit ("mpk_halo_level") or automatic: the
        //      deepest level the supplied rows cover whose ghost count does not exceed the owned rows (redundant work <= 2x),
        //      minimised over the ranks so that everybody follows the same exchange schedule.
        const int want = ctx->opt_mpk_halo_level > 0 ? (int)std::min<int64_t>(ctx->opt_mpk_halo_level, s_max) : s_max;
        int covered = want;
        int st = level_sets_host(ctx, n_glob, row_begin, row_end, rowptr, colind, lo, hi, want, level, nullptr,
                                 ctx->opt_mpk_halo_level > 0 ? nullptr : &covered);
        if (st != CALZ_OK) {
            delete m;
            return st;
        }
        int L = covered;
        if (ctx->opt_mpk_halo_level <= 0) {
            std::vector<int64_t> cnt(want + 2, 0);
            for (int64_t j = 0; j < n_glob; ++j)
                if (level[j] > 0) cnt[level[j]]++;
            int64_t ghosts = 0;
            int best = 1;
            for (int k = 1; k <= covered; ++k) {
                ghosts += cnt[k];
                if (ghosts <= m->n_own) best = k;
            }
            L = std::max(1, std::min(best, covered));
            if (covered < 1) {
                delete m;
                return set_error(ctx, CALZ_ERR_CLOSURE, "the owned rows [%lld,%lld) were not all supplied", (long long)lo, (long long)hi);
            }
        }
        {   // min over the ranks (a P-vector of proposals through the small all-reduce)
            std::vector<double> prop(P, 0.0);
            prop[me] = (double)L;
            double* d_prop = nullptr;
            CALZ_CUDA(ctx, cudaMalloc(&d_prop, P * sizeof(double)));
            CALZ_CUDA(ctx, cudaMemcpy(d_prop, prop.data(), P * sizeof(double), cudaMemcpyHostToDevice));
            CALZ_TRY(allreduce_sum(ctx, d_prop, P));
            CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            CALZ_CUDA(ctx, cudaMemcpy(prop.data(), d_prop, P * sizeof(double), cudaMemcpyDeviceToHost));
            cudaFree(d_prop);
            for (int q = 0; q < P; ++q) L = std::min(L, (int)prop[q]);
        }
        if (L < 1) {
            delete m;
            return set_error(ctx, CALZ_ERR_CLOSURE, "no common ghost-closure depth");
        }
        for (int64_t j = 0; j < n_glob; ++j)
            if (level[j] > L) level[j] = (int8_t)-1;
        m->halo_level = L;
        for (int64_t j = 0; j < lo; ++j)
            if (level[j] > 0) m->ghost_glob.push_back(j);
        m->own_off = (int64_t)m->ghost_glob.size();
        for (int64_t j = hi; j < n_glob; ++j)
            if (level[j] > 0) m->ghost_glob.push_back(j);
        m->n_loc = m->n_own + (int64_t)m->ghost_glob.size();
    }
    const int64_t n_loc = m->n_loc, own_off = m->own_off;
    const int64_t n_ghost = n_loc - m->n_own;
    auto loc_to_glob = [&](int64_t l) -> int64_t {
        if (l < own_off) return m->ghost_glob[l];
        if (l < own_off + m->n_own) return lo + (l - own_off);
        return m->ghost_glob[l - m->n_own];
    };
    auto glob_to_loc = [&](int64_t g) -> int64_t {      // -1 if not in R_s
        if (g >= lo && g < hi) return own_off + (g - lo);
        if (P == 1) return -1;
        if (level[g] < 0) return -1;
        if (g < lo) return std::lower_bound(m->ghost_glob.begin(), m->ghost_glob.begin() + own_off, g) - m->ghost_glob.begin();
        return m->n_own + (std::lower_bound(m->ghost_glob.begin() + own_off, m->ghost_glob.end(), g) - m->ghost_glob.begin());
    };

    // ---- local CSR: rows with level <= L-1 (level-L rows are only read, never computed).  Rows with more than kLongRow
    //      entries go to the long-row structure instead (main row length 0).
    const int HL = m->halo_level;
    constexpr int64_t kLongRow = 2048, kLongSeg = 4096;
    std::vector<int32_t> long_row, long_seg0(1, 0), long_segptr(1, 0), long_segrow;
    std::vector<int64_t> long_len;
    std::vector<int64_t> h_rowptr64(n_loc + 1, 0);
    for (int64_t l = 0; l < n_loc; ++l) {
        int64_t g = loc_to_glob(l);
        int lev = (P == 1) ? 0 : level[g];
        int64_t cnt = 0;
        if (lev <= HL - 1) {
            if (g < row_begin || g >= row_end) {
                delete m;
                return set_error(ctx, CALZ_ERR_CLOSURE, "row %lld (level %d) needed but not supplied", (long long)g, lev);
            }
            if (P == 1) cnt = rowptr[g - row_begin + 1] - rowptr[g - row_begin];
            else
                for (int64_t e = rowptr[g - row_begin]; e < rowptr[g - row_begin + 1]; ++e)
                    if (level[colind[e]] >= 0) ++cnt;
        }
        if (cnt > kLongRow && n_loc < (int64_t)2147483647) {
            long_row.push_back((int32_t)l);
            long_len.push_back(cnt);
            for (int64_t o = 0; o < cnt; o += kLongSeg) {
                long_segrow.push_back((int32_t)long_row.size() - 1);
                long_segptr.push_back(long_segptr.back() + (int32_t)std::min<int64_t>(kLongSeg, cnt - o));
            }
            long_seg0.push_back((int32_t)long_segrow.size());
            cnt = 0;
        }
        h_rowptr64[l + 1] = h_rowptr64[l] + cnt;
    }
    m->n_long = (int64_t)long_row.size();
    m->n_long_seg = (int64_t)long_segrow.size();
    const int64_t nnz_long = long_segptr.back();
    m->nnz_loc = h_rowptr64[n_loc] + nnz_long;
    if (m->nnz_loc >= (int64_t)2147483647 || n_loc >= (int64_t)2147483647) {
        delete m;
        return set_error(ctx, CALZ_ERR_UNSUPPORTED, "local matrix too large for 32-bit indices (nnz_loc=%lld)", (long long)m->nnz_loc);
    }
    std::vector<int32_t> h_rowptr(n_loc + 1);
    for (int64_t l = 0; l <= n_loc; ++l) h_rowptr[l] = (int32_t)h_rowptr64[l];
    std::vector<int32_t> h_col((size_t)(m->nnz_loc - nnz_long));
    std::vector<double> h_val((size_t)(m->nnz_loc - nnz_long));
    std::vector<int32_t> lg_col((size_t)nnz_long);
    std::vector<double> lg_val((size_t)nnz_long);
    int64_t bw = 0;
    {   // long rows first (their main row length is 0)
        for (size_t r = 0; r < long_row.size(); ++r) {
            const int64_t l = long_row[r], g = loc_to_glob(l);
            int64_t w = long_segptr[long_seg0[r]];
            for (int64_t e = rowptr[g - row_begin]; e < rowptr[g - row_begin + 1]; ++e) {
                int64_t jl = (P == 1) ? (int64_t)colind[e] : glob_to_loc(colind[e]);
                if (jl < 0) continue;
                lg_col[w] = (int32_t)jl;
                lg_val[w] = val[e];
                bw = std::max<int64_t>(bw, jl > l ? jl - l : l - jl);
                ++w;
            }
        }
    }
    for (int64_t l = 0; l < n_loc; ++l) {
        if (h_rowptr[l + 1] == h_rowptr[l]) continue;
        int64_t g = loc_to_glob(l);
        int64_t w = h_rowptr[l];
        for (int64_t e = rowptr[g - row_begin]; e < rowptr[g - row_begin + 1]; ++e) {
            int64_t jl = (P == 1) ? (int64_t)colind[e] : glob_to_loc(colind[e]);
            if (jl < 0) continue;
            h_col[w] = (int32_t)jl;
            h_val[w] = val[e];
            bw = std::max<int64_t>(bw, jl > l ? jl - l : l - jl);
            ++w;
        }
    }
    m->bandwidth = bw;

    // ---- hull of the rows with level <= L (the active range of MPK step k of an s-step call is L = s-k)
    m->hull_lo.assign(s_max + 1, 0);
    m->hull_hi.assign(s_max + 1, n_loc);
    if (P > 1) {
        for (int L = 0; L <= HL; ++L) {
            int64_t a = own_off, b = own_off + m->n_own;
            for (int64_t l = 0; l < own_off; ++l)
                if (level[m->ghost_glob[l]] <= L) { a = l; break; }
            for (int64_t l = n_loc - 1; l >= own_off + m->n_own; --l)
                if (level[m->ghost_glob[l - m->n_own]] <= L) { b = l + 1; break; }
            m->hull_lo[L] = a;
            m->hull_hi[L] = b;
        }
    }

    // ---- exchange plan
    m->recv_off.assign(P, 0); m->recv_cnt.assign(P, 0);
    m->recv_glob.assign(P, {}); m->send_glob.assign(P, {});
    m->send_off.assign(P, 0); m->send_cnt.assign(P, 0); m->send_contig.assign(P, 0);
    if (P > 1) {
        for (int q = 0; q < P; ++q) {
            if (q == me) continue;
            auto b0 = std::lower_bound(m->ghost_glob.begin(), m->ghost_glob.end(), bounds[q]);
            auto b1 = std::lower_bound(m->ghost_glob.begin(), m->ghost_glob.end(), bounds[q + 1]);
            m->recv_glob[q].assign(b0, b1);
            m->recv_cnt[q] = b1 - b0;
            int64_t gi = b0 - m->ghost_glob.begin();                 // index into ghost list
            m->recv_off[q] = gi < own_off ? gi : gi + m->n_own;      // local index of the run
        }
        // tell every peer what we need: counts through an all-reduced P x P table, lists through send/recv
        // (indices < 2^53 travel exactly as fp64)
        size_t total_recv = 0;
        for (int q = 0; q < P; ++q) total_recv += m->recv_cnt[q];
        std::vector<double> tab((size_t)P * P, 0.0);
        for (int q = 0; q < P; ++q) tab[(size_t)me * P + q] = (double)m->recv_cnt[q];
        double* d_tab = nullptr;
        CALZ_CUDA(ctx, cudaMalloc(&d_tab, tab.size() * sizeof(double)));
        CALZ_CUDA(ctx, cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
        CALZ_TRY(allreduce_sum(ctx, d_tab, tab.size()));
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CALZ_CUDA(ctx, cudaMemcpy(tab.data(), d_tab, tab.size() * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(d_tab);
        size_t total_send = 0;
        for (int q = 0; q < P; ++q) {
            m->send_cnt[q] = (q == me) ? 0 : (int64_t)tab[(size_t)q * P + me];   // q needs this many from me
            m->send_off[q] = (int64_t)total_send;
            total_send += m->send_cnt[q];
        }
        std::vector<double> want(std::max<size_t>(total_recv, 1)), give(std::max<size_t>(total_send, 1));
        {
            size_t w = 0;
            for (int q = 0; q < P; ++q)
                for (int64_t g : m->recv_glob[q]) want[w++] = (double)g;
        }
        double *d_want = nullptr, *d_give = nullptr;
        CALZ_CUDA(ctx, cudaMalloc(&d_want, want.size() * sizeof(double)));
        CALZ_CUDA(ctx, cudaMalloc(&d_give, give.size() * sizeof(double)));
        CALZ_CUDA(ctx, cudaMemcpy(d_want, want.data(), want.size() * sizeof(double), cudaMemcpyHostToDevice));
        CALZ_NCCL(ctx, ctx->nccl->GroupStart());
        {
            size_t w = 0;
            for (int q = 0; q < P; ++q) {
                if (m->recv_cnt[q]) CALZ_NCCL(ctx, ctx->nccl->Send(d_want + w, m->recv_cnt[q], ncclFloat64, q, ctx->comm, ctx->stream));
                w += m->recv_cnt[q];
                if (m->send_cnt[q]) CALZ_NCCL(ctx, ctx->nccl->Recv(d_give + m->send_off[q], m->send_cnt[q], ncclFloat64, q, ctx->comm, ctx->stream));
            }
        }
        CALZ_NCCL(ctx, ctx->nccl->GroupEnd());
        CALZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CALZ_CUDA(ctx, cudaMemcpy(give.data(), d_give, give.size() * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(d_want);
        cudaFree(d_give);
        std::vector<int32_t> send_idx(std::max<size_t>(total_send, 1));
        for (int q = 0; q < P; ++q) {
            m->send_glob[q].resize(m->send_cnt[q]);
            bool contig = true;
            for (int64_t i = 0; i < m->send_cnt[q]; ++i) {
                int64_t g = (int64_t)give[m->send_off[q] + i];
                if (g < lo || g >= hi) {
                    delete m;
                    return set_error(ctx, CALZ_ERR_CLOSURE, "peer %d asked for row %lld which rank %d does not own", q, (long long)g, me);
                }
                m->send_glob[q][i] = g;
                send_idx[m->send_off[q] + i] = (int32_t)(own_off + (g - lo));
                if (i > 0 && g != m->send_glob[q][i - 1] + 1) contig = false;
            }
            m->send_contig[q] = contig ? 1 : 0;
        }
        CALZ_TRY(upload(ctx, &m->d_send_idx, send_idx));
        CALZ_CUDA(ctx, cudaMalloc(&m->d_send_buf, std::max<size_t>(total_send, 1) * sizeof(double)));
    }
    (void)n_ghost;

    // ---- layout decision
    const int C = 32;
    m->sell_slices = (n_loc + C - 1) / C;
    std::vector<int32_t> rowlen(n_loc);
    for (int64_t l = 0; l < n_loc; ++l) rowlen[l] = h_rowptr[l + 1] - h_rowptr[l];
    auto padded_for = [&](int sigma, std::vector<int32_t>* perm) -> int64_t {
        std::vector<int32_t> order;
        const std::vector<int32_t>* len = &rowlen;
        std::vector<int32_t> sorted_len;
        if (sigma > C) {
            order.resize(n_loc);
            std::iota(order.begin(), order.end(), 0);
            for (int64_t w0 = 0; w0 < n_loc; w0 += sigma) {
                int64_t w1 = std::min<int64_t>(n_loc, w0 + sigma);
                std::stable_sort(order.begin() + w0, order.begin() + w1,
                                 [&](int32_t a, int32_t b) { return rowlen[a] > rowlen[b]; });
            }
            sorted_len.resize(n_loc);
            for (int64_t l = 0; l < n_loc; ++l) sorted_len[l] = rowlen[order[l]];
            len = &sorted_len;
        }
        int64_t tot = 0;
        for (int64_t s0 = 0; s0 < n_loc; s0 += C) {
            int32_t w = 0;
            for (int64_t l = s0; l < std::min<int64_t>(n_loc, s0 + C); ++l) w = std::max(w, (*len)[l]);
            tot += (int64_t)w * C;
        }
        if (perm) perm->swap(order);
        return tot;
    };
    int sigma = C;
    std::vector<int32_t> perm;
    int64_t padded = padded_for(C, nullptr);
    if (ctx->opt_sell_sigma > C) {
        sigma = (int)round_up(ctx->opt_sell_sigma, C);
        padded = padded_for(sigma, &perm);
    } else if (ctx->opt_sell_sigma == 0 && m->nnz_loc > 0 && (double)padded > 1.10 * (double)m->nnz_loc) {
        std::vector<int32_t> p2;
        int64_t pad2 = padded_for(4096, &p2);
        if ((double)pad2 < 0.9 * (double)padded) {
            sigma = 4096;
            padded = pad2;
            perm.swap(p2);
        }
    }
    // ---- dictionary-coded SELL: one byte per non-zero when the matrix has <= 255 distinct (column offset, value) pairs
    //      (every constant-coefficient stencil; lossless, same arithmetic order as CSR/SELL)
    std::vector<int32_t> dict_off;
    std::vector<double> dict_val;
    std::vector<uint8_t> codes;          // per CSR entry
    bool dict_ok = (layout == CALZ_LAYOUT_AUTO || layout == CALZ_LAYOUT_SELL_DICT) && m->nnz_loc > 0;
    if (dict_ok) {
        std::map<std::pair<int32_t, uint64_t>, int> lut;
        codes.resize((size_t)m->nnz_loc);
        for (int64_t l = 0; l < n_loc && dict_ok; ++l)
            for (int32_t e = h_rowptr[l]; e < h_rowptr[l + 1]; ++e) {
                uint64_t bits;
                memcpy(&bits, &h_val[e], 8);
                auto key = std::make_pair((int32_t)(h_col[e] - (int32_t)l), bits);
                auto it = lut.find(key);
                if (it == lut.end()) {
                    if (lut.size() >= 255) { dict_ok = false; break; }
                    it = lut.emplace(key, (int)lut.size()).first;
                    dict_off.push_back(key.first);
                    dict_val.push_back(h_val[e]);
                }
                codes[e] = (uint8_t)it->second;
            }
        int32_t wmax = 0;
        for (int64_t l = 0; l < n_loc; ++l) wmax = std::max(wmax, rowlen[l]);
        if (wmax > 64) dict_ok = false;                       // very long rows: the byte stream per lane gets too ragged
        if (n_loc >= ((int64_t)1 << 28) - 64) dict_ok = false;   // byte offsets and row indices are 32-bit in the kernel
    }
    if (layout == CALZ_LAYOUT_SELL_DICT && !dict_ok) {
        delete m;
        return set_error(ctx, CALZ_ERR_UNSUPPORTED, "SELL_DICT layout needs <= 255 distinct (offset, value) pairs and rows <= 64 entries");
    }
    if (layout == CALZ_LAYOUT_AUTO && dict_ok && ctx->opt_sell_dict) layout = CALZ_LAYOUT_SELL_DICT;
    if (layout == CALZ_LAYOUT_AUTO)
        layout = (m->nnz_loc > 0 && (double)padded <= 1.30 * (double)m->nnz_loc && padded < (int64_t)2147483647)
                     ? CALZ_LAYOUT_SELL : CALZ_LAYOUT_CSR;
    if (layout == CALZ_LAYOUT_SELL && padded >= (int64_t)64 * 2147483647)
        layout = CALZ_LAYOUT_CSR;
    m->layout = layout;

    if (layout == CALZ_LAYOUT_SELL_DICT) {
        // per slice of 32 rows: ceil(width/8) blocks of 32 lanes x 8 code bytes (one coalesced 256-B load per block)
        std::vector<int32_t> slice_ptr(m->sell_slices + 2, 0);                  // + one padding entry (the kernel reads [s .. s+2])
        int64_t blocks = 0;
        for (int64_t sl = 0; sl < m->sell_slices; ++sl) {
            int32_t w = 0;
            for (int64_t r = sl * C; r < std::min<int64_t>(n_loc, sl * C + C); ++r) w = std::max(w, rowlen[r]);
            slice_ptr[sl] = (int32_t)blocks;
            blocks += (w + 7) / 8;
        }
        slice_ptr[m->sell_slices] = (int32_t)blocks;
        slice_ptr[m->sell_slices + 1] = (int32_t)blocks;
        std::vector<uint8_t> packed((size_t)blocks * 256, (uint8_t)255);        // 255 = padding code
        for (int64_t sl = 0; sl < m->sell_slices; ++sl)
            for (int64_t r = sl * C; r < std::min<int64_t>(n_loc, sl * C + C); ++r)
                for (int32_t j = 0; j < rowlen[r]; ++j)
                    packed[((size_t)(slice_ptr[sl] + j / 8) * 32 + (size_t)(r - sl * C)) * 8 + (j % 8)] = codes[h_rowptr[r] + j];
        // how often do the 32 lanes of a (block, slot) hold ONE code?  (sampled)  A warp-uniform code is a broadcast read
        // from the constant bank; divergent codes would serialise there, so ragged matrices keep the dictionary in shared memory
        {
            int64_t uni = 0, tot = 0;
            const int64_t step = std::max<int64_t>(1, blocks / 4096) | 1;     // odd: visits every residue of a power-of-two period
            for (int64_t b = 0; b < blocks; b += step)
                for (int q = 0; q < 8; ++q) {
                    const uint8_t* pb = packed.data() + (size_t)b * 256 + q;
                    bool same = true;
                    for (int l = 1; l < 32 && same; ++l) same = pb[l * 8] == pb[0];
                    uni += same;
                    ++tot;
                }
            m->dict_uniform = tot ? (double)uni / (double)tot : 0.0;
        }
        std::vector<double> dict(2 * 256, 0.0);                                   // {value, offset-as-int64 bits} pairs, 16 B each
        for (size_t k = 0; k < dict_val.size(); ++k) {
            dict[2 * k] = dict_val[k];
            long long o = dict_off[k];
            memcpy(&dict[2 * k + 1], &o, 8);
            m->h_dict->e[k].v = dict_val[k];
            m->h_dict->e[k].offb = (int)(o * 8);
        }
        // ---- slice patterns (see matrix.h): one-block slices of 32 complete rows whose lanes agree on the value of every offset
        {
            std::vector<uint8_t> spat((size_t)m->sell_slices + 4, (uint8_t)255);
            struct Pat { int cnt; int32_t off[8]; uint64_t vb[8]; uint32_t mask[8]; int64_t freq; };
            std::vector<Pat> pats;
            std::map<std::vector<uint64_t>, int> pat_lut;
            std::vector<int32_t> prov((size_t)m->sell_slices, -1);                 // provisional pattern number per slice
            for (int64_t sl = 0; sl < m->sell_slices; ++sl) {
                if (slice_ptr[sl + 1] - slice_ptr[sl] > 1 || (sl + 1) * C > n_loc) continue;
                Pat P{};
                bool ok = true;
                if (slice_ptr[sl + 1] > slice_ptr[sl]) {
                    const uint8_t* pb = packed.data() + (size_t)slice_ptr[sl] * 256;
                    for (int lane = 0; lane < 32 && ok; ++lane)
                        for (int q = 0; q < 8 && ok; ++q) {
                            const uint8_t cd = pb[lane * 8 + q];
                            if (cd == 255) continue;
                            const int32_t o = dict_off[cd];
                            uint64_t vb;
                            memcpy(&vb, &dict_val[cd], 8);
                            int k = 0;
                            while (k < P.cnt && P.off[k] != o) ++k;
                            if (k == P.cnt) {
                                if (P.cnt == 8) { ok = false; break; }
                                P.off[k] = o; P.vb[k] = vb; P.mask[k] = 0; ++P.cnt;
                            } else if (P.vb[k] != vb) { ok = false; break; }
                            if (P.mask[k] & (1u << lane)) { ok = false; break; }      // the same offset twice in one row
                            P.mask[k] |= 1u << lane;
                        }
                }                                                                      // else: 32 empty rows, the empty pattern
                if (!ok) continue;
                std::vector<uint64_t> key;
                for (int k = 0; k < P.cnt; ++k) { key.push_back((uint64_t)(uint32_t)P.off[k]); key.push_back(P.vb[k]); key.push_back(P.mask[k]); }
                std::sort(key.begin(), key.end());      // (order-insensitive key; entries are re-sorted by offset below)
                auto it = pat_lut.find(key);
                if (it == pat_lut.end()) {
                    if (pats.size() >= 4096) continue;
                    it = pat_lut.emplace(key, (int)pats.size()).first;
                    pats.push_back(P);
                }
                pats[it->second].freq++;
                prov[sl] = it->second;
            }
            // pattern 0 = the most frequent one, entries in ascending offset order (= ascending column = the CSR order of every
            // row); every other pattern must be a sub-pattern of it
            int64_t covered = 0;
            m->n_pat = 0;
            if (!pats.empty()) {
                int top = 0;
                for (size_t q = 1; q < pats.size(); ++q)
                    if (pats[q].freq > pats[top].freq) top = (int)q;
                const Pat& P0 = pats[top];
                int ord[8];
                for (int k = 0; k < P0.cnt; ++k) ord[k] = k;
                std::sort(ord, ord + P0.cnt, [&](int a, int b) { return P0.off[a] < P0.off[b]; });
                for (int k = 0; k < P0.cnt; ++k) {
                    memcpy(&m->h_pat->e0[k].v, &P0.vb[ord[k]], 8);
                    m->h_pat->e0[k].offb = P0.off[ord[k]] * 8;
                    m->h_pat->e0[k].mask = 0xffffffffu;
                }
                std::vector<int> newid(pats.size(), -1);
                auto assign = [&](int q) -> bool {                      // masks of pattern q on the entries of pattern 0
                    if (m->n_pat >= 32) return false;
                    unsigned mk[8] = {};
                    for (int a = 0; a < pats[q].cnt; ++a) {
                        int k = 0;
                        while (k < P0.cnt && !(P0.off[ord[k]] == pats[q].off[a] && P0.vb[ord[k]] == pats[q].vb[a])) ++k;
                        if (k == P0.cnt) return false;
                        mk[k] = pats[q].mask[a];
                    }
                    for (int k = 0; k < 8; ++k) m->h_pat->mask[m->n_pat][k] = mk[k];
                    newid[q] = m->n_pat++;
                    return true;
                };
                assign(top);
                bool full = P0.cnt >= 1;
                for (int k = 0; k < P0.cnt; ++k) full &= P0.mask[k] == 0xffffffffu;
                m->pat_cnt0 = full ? P0.cnt : 0;
                std::vector<int> byfreq(pats.size());
                std::iota(byfreq.begin(), byfreq.end(), 0);
                std::stable_sort(byfreq.begin(), byfreq.end(), [&](int a, int b) { return pats[a].freq > pats[b].freq; });
                for (int q : byfreq)
                    if (q != top) assign(q);
                for (int64_t sl = 0; sl < m->sell_slices; ++sl)
                    if (prov[sl] >= 0 && newid[prov[sl]] >= 0) { spat[sl] = (uint8_t)newid[prov[sl]]; ++covered; }
            }
            m->pat_cover = m->sell_slices ? (double)covered / (double)m->sell_slices : 0.0;
            {   // pairing phase: the one with more all-interior pairs (both slices of the item carry pattern 0)
                int64_t cnt[2] = {0, 0};
                for (int ph = 0; ph < 2; ++ph)
                    for (int64_t sl = -ph; sl + 1 < m->sell_slices; sl += 2)
                        if (sl >= 0 && spat[sl] == 0 && spat[sl + 1] == 0) ++cnt[ph];
                m->pat_phase = cnt[1] > cnt[0] ? 1 : 0;
            }
            CALZ_TRY(upload(ctx, &m->d_slice_pat, spat));
        }
        m->sell_padded = blocks * 256;
        m->dict_size = (int)dict_val.size();
        CALZ_TRY(upload(ctx, &m->d_slice_ptr, slice_ptr));
        CALZ_TRY(upload(ctx, &m->d_codes, packed));
        CALZ_TRY(upload(ctx, &m->d_dict, dict));
        // x staging plan: for a CTA owning rows [r0, r0+R) the entries with offset o read x[r0+o .. r0+R+o); offsets whose
        // ranges overlap are merged into one segment (one TMA bulk copy each)
        for (int R : {2048, 1024, 512}) {
            if (ctx->opt_mpk_xs_rows > 0 && R > ctx->opt_mpk_xs_rows) continue;
            std::vector<int32_t> offs(dict_off);
            std::sort(offs.begin(), offs.end());
            offs.erase(std::unique(offs.begin(), offs.end()), offs.end());
            int ng = 0, total = 0, omin[8], omax[8];
            bool ok = !offs.empty();
            for (size_t k = 0; k < offs.size() && ok; ++k) {
                if (ng > 0 && (int64_t)offs[k] - omax[ng - 1] < R) omax[ng - 1] = offs[k];
                else if (ng < 8) { omin[ng] = offs[k]; omax[ng] = offs[k]; ++ng; }
                else ok = false;
            }
            int base[8], len[8], amin[8];
            for (int g = 0; g < ng && ok; ++g) {
                amin[g] = omin[g] - (((omin[g] % 2) + 2) % 2);           // even, <= omin
                len[g] = R + (omax[g] - amin[g]) + 2;
                len[g] += len[g] & 1;
                base[g] = total;
                total += len[g];
            }
            if (!ok || (size_t)total * 8 > 96 * 1024 || n_loc < 4 * R) continue;
            m->xs_rows = R; m->xs_groups = ng; m->xs_total = total;
            std::vector<int> xoff(256, 0);
            for (int g = 0; g < ng; ++g) { m->xs_omin[g] = amin[g]; m->xs_len[g] = len[g]; m->xs_base[g] = base[g]; }
            for (size_t k = 0; k < dict_off.size(); ++k)
                for (int g = 0; g < ng; ++g)
                    if (dict_off[k] >= omin[g] && dict_off[k] <= omax[g]) xoff[k] = base[g] + (dict_off[k] - amin[g]);
            CALZ_TRY(upload(ctx, &m->d_xs_off, xoff));
            break;
        }
    } else if (layout == CALZ_LAYOUT_CSR) {
        CALZ_TRY(upload(ctx, &m->d_rowptr, h_rowptr));
        CALZ_TRY(upload(ctx, &m->d_colind, h_col));
        CALZ_TRY(upload(ctx, &m->d_val, h_val));
        double mean = n_loc ? (double)m->nnz_loc / (double)n_loc : 1.0;
        int lanes = 1;
        while (lanes < 32 && lanes * 2 <= mean + 1.0) lanes *= 2;
        if (ctx->opt_csr_lanes > 0) lanes = (int)ctx->opt_csr_lanes;
        m->csr_lanes = lanes;
    } else {
        m->sell_sigma = sigma;
        m->sell_padded = padded;
        std::vector<int32_t> slice_ptr(m->sell_slices + 1, 0);
        std::vector<int32_t> s_col((size_t)padded);
        std::vector<double> s_val((size_t)padded, 0.0);
        int64_t off = 0;   // in units of 32 entries
        for (int64_t sl = 0; sl < m->sell_slices; ++sl) {
            int64_t r0 = sl * C, r1 = std::min<int64_t>(n_loc, r0 + C);
            int32_t w = 0;
            for (int64_t r = r0; r < r1; ++r) {
                int64_t l = perm.empty() ? r : perm[r];
                w = std::max(w, rowlen[l]);
            }
            slice_ptr[sl] = (int32_t)off;
            for (int64_t r = r0; r < r0 + C; ++r) {
                int64_t l = (r < r1) ? (perm.empty() ? r : (int64_t)perm[r]) : -1;
                int32_t len = l >= 0 ? rowlen[l] : 0;
                int32_t self = (int32_t)(l >= 0 ? l : 0);
                for (int32_t j = 0; j < w; ++j) {
                    size_t dst = (size_t)(off + j) * C + (r - r0);
                    if (j < len) {
                        s_col[dst] = h_col[h_rowptr[l] + j];
                        s_val[dst] = h_val[h_rowptr[l] + j];
                    } else {
                        s_col[dst] = self;     // padding: zero value, harmless in-range column
                    }
                }
            }
            off += w;
        }
        slice_ptr[m->sell_slices] = (int32_t)off;
        CALZ_TRY(upload(ctx, &m->d_slice_ptr, slice_ptr));
        CALZ_TRY(upload(ctx, &m->d_sell_col, s_col));
        CALZ_TRY(upload(ctx, &m->d_sell_val, s_val));
        if (!perm.empty()) CALZ_TRY(upload(ctx, &m->d_perm, perm));
    }

    if (m->n_long) {
        CALZ_TRY(upload(ctx, &m->d_long_row, long_row));
        CALZ_TRY(upload(ctx, &m->d_long_seg0, long_seg0));
        CALZ_TRY(upload(ctx, &m->d_long_segptr, long_segptr));
        CALZ_TRY(upload(ctx, &m->d_long_segrow, long_segrow));
        CALZ_TRY(upload(ctx, &m->d_long_col, lg_col));
        CALZ_TRY(upload(ctx, &m->d_long_val, lg_val));
        CALZ_CUDA(ctx, cudaMalloc(&m->d_long_part, (size_t)m->n_long_seg * sizeof(double)));
    }

    // ---- basis workspace
    m->ldW = round_up(n_loc, 32);
    // The OWNED rows of every column must start 16-byte aligned on every rank (TMA bulk copies; and every rank has to take the
    // same orthogonalisation path, or the collectives no longer match): shift the workspace by one double when own_off is odd.
    m->W_pad = (int)(m->own_off & 1);
    size_t wbytes = ((size_t)m->ldW * (size_t)(s_max + 1) + 2) * sizeof(double);
    cudaError_t e = cudaMalloc(&m->d_W_alloc, wbytes);
    if (e != cudaSuccess) {
        calz_mat_destroy(m);
        return set_error(ctx, CALZ_ERR_ALLOC, "basis workspace %zu bytes: %s", wbytes, cudaGetErrorString(e));
    }
    CALZ_CUDA(ctx, cudaMemset(m->d_W_alloc, 0, wbytes));
    m->d_W = m->d_W_alloc + m->W_pad;
    if (P > 1) CALZ_TRY(p2p_halo_setup(m));
    *out = m;
    return CALZ_OK;
}

int calz_mat_create_csc64(calz_ctx* ctx, int64_t n, const uint64_t* jc, const uint64_t* ir, const double* pr,
                          int s_max, int layout, calz_mat** out) {
    if (!ctx || !jc || !ir || !pr || n <= 0) return set_error(ctx, CALZ_ERR_BADARG, "calz_mat_create_csc64: bad arguments");
    if (n >= (int64_t)2147483647) return set_error(ctx, CALZ_ERR_UNSUPPORTED, "n too large for 32-bit column indices");
    // CSC(A) -> CSR(A) by a counting transpose (A need not be symmetric); columns come out ascending.
    const int64_t nnz = (int64_t)jc[n];
    std::vector<int64_t> rowptr(n + 1, 0);
    for (int64_t e = 0; e < nnz; ++e) rowptr[ir[e] + 1]++;
    for (int64_t i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int32_t> col((size_t)nnz);
    std::vector<double> val((size_t)nnz);
    std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
    for (int64_t j = 0; j < n; ++j)
        for (uint64_t e = jc[j]; e < jc[j + 1]; ++e) {
            int64_t w = fill[ir[e]]++;
            col[w] = (int32_t)j;
            val[w] = pr[e];
        }
    return calz_mat_create_csr(ctx, n, 0, n, rowptr.data(), col.data(), val.data(), s_max, layout, out);
}

int calz_mat_info(const calz_mat* m, const char* what, int64_t* value) {
    if (!m || !what || !value) return CALZ_ERR_BADARG;
    if (!strcmp(what, "n_glob")) *value = m->n_glob;
    else if (!strcmp(what, "n_own")) *value = m->n_own;
    else if (!strcmp(what, "n_loc")) *value = m->n_loc;
    else if (!strcmp(what, "own_off")) *value = m->own_off;
    else if (!strcmp(what, "row_lo")) *value = m->row_lo;
    else if (!strcmp(what, "row_hi")) *value = m->row_hi;
    else if (!strcmp(what, "nnz_loc")) *value = m->nnz_loc;
    else if (!strcmp(what, "layout")) *value = m->layout;
    else if (!strcmp(what, "sell_padded_nnz")) *value = m->sell_padded;
    else if (!strcmp(what, "sell_sigma")) *value = m->sell_sigma;
    else if (!strcmp(what, "csr_lanes")) *value = m->csr_lanes;
    else if (!strcmp(what, "n_ghost")) *value = m->n_loc - m->n_own;
    else if (!strcmp(what, "bandwidth")) *value = m->bandwidth;
    else if (!strcmp(what, "s_max")) *value = m->s_max;
    else if (!strcmp(what, "halo_level")) *value = m->halo_level;
    else if (!strcmp(what, "n_long_rows")) *value = m->n_long;
    else if (!strcmp(what, "ldW")) *value = m->ldW;
    else if (!strcmp(what, "dict_size")) *value = m->dict_size;
    else if (!strcmp(what, "dict_uniform_pct")) *value = (int64_t)(100.0 * m->dict_uniform);
    else if (!strcmp(what, "n_patterns")) *value = m->n_pat;
    else if (!strcmp(what, "pattern_pair_phase")) *value = m->pat_phase;
    else if (!strcmp(what, "pattern_cover_pct")) *value = (int64_t)(100.0 * m->pat_cover);
    else if (!strcmp(what, "xs_rows")) *value = m->xs_rows;
    else if (!strcmp(what, "xs_groups")) *value = m->xs_groups;
    else if (!strcmp(what, "p2p_halo")) *value = m->p2p_halo ? 1 : 0;
    else if (!strcmp(what, "p2p_allreduce")) *value = (m->ctx && m->ctx->p2p.enabled) ? 1 : 0;
    else return set_error(m->ctx, CALZ_ERR_BADARG, "calz_mat_info: unknown key '%s'", what);
    return CALZ_OK;
}

static int copy_list(const std::vector<int64_t>& v, int64_t* idx_out, int64_t* count) {
    if (!count) return CALZ_ERR_BADARG;
    *count = (int64_t)v.size();
    if (idx_out) memcpy(idx_out, v.data(), v.size() * sizeof(int64_t));
    return CALZ_OK;
}

int calz_mat_ghost_indices(const calz_mat* m, int64_t* idx_out, int64_t* count) {
    if (!m) return CALZ_ERR_BADARG;
    return copy_list(m->ghost_glob, idx_out, count);
}

int calz_mat_recv_list(const calz_mat* m, int peer, int64_t* idx_out, int64_t* count) {
    if (!m || peer < 0 || peer >= (int)m->recv_glob.size()) return CALZ_ERR_BADARG;
    return copy_list(m->recv_glob[peer], idx_out, count);
}

int calz_mat_send_list(const calz_mat* m, int peer, int64_t* idx_out, int64_t* count) {
    if (!m || peer < 0 || peer >= (int)m->send_glob.size()) return CALZ_ERR_BADARG;
    return copy_list(m->send_glob[peer], idx_out, count);
}

}  // extern "C"
