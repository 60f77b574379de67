// Device sparse matrix of one rank: owned rows + level-s ghost closure, CSR and/or SELL-32-sigma.
#pragma once
#include "calz_internal.h"

struct calz_mat {
    calz_ctx* ctx = nullptr;
    int64_t n_glob = 0, row_lo = 0, row_hi = 0;      // owned global rows [row_lo,row_hi)
    int64_t n_own = 0, n_loc = 0, own_off = 0;       // local index space: ascending global index over R_s
    int64_t nnz_loc = 0, bandwidth = 0;
    int s_max = 0;                                   // columns of the basis workspace - 1
    int halo_level = 0;                              // L: depth of the ghost closure = MPK steps per halo exchange (L = s_max is PA1:
                                                     // one exchange per block; L = 1 is the classic per-step exchange, what a power-law
                                                     // graph whose level-1 closure is already every row needs)
    int layout = CALZ_LAYOUT_CSR;

    std::vector<int64_t> ghost_glob;                 // sorted global indices of ghost_s(p)
    std::vector<int64_t> hull_lo, hull_hi;           // hull of local rows with level <= L, L = 0..s_max

    // exchange plan (per peer)
    std::vector<int64_t> recv_off, recv_cnt;         // ghosts from peer q: contiguous run of local indices
    std::vector<std::vector<int64_t>> recv_glob;     // ... and their global indices (bit-exact object)
    std::vector<std::vector<int64_t>> send_glob;     // what peer q wants from me (global indices, sorted)
    std::vector<int64_t> send_off, send_cnt;         // offsets into d_send_idx / d_send_buf
    std::vector<char> send_contig;                   // send list is a contiguous local range: no pack
    int32_t* d_send_idx = nullptr;                   // local indices to pack, all peers concatenated
    double* d_send_buf = nullptr;

    // peer-memory halo (p2p.cu): the peers' basis workspaces and where my rows land in them
    bool p2p_halo = false;
    std::vector<double*> peer_W;
    std::vector<void*> peer_W_base;                  // what cudaIpcOpenMemHandle returned (to close it)
    std::vector<long long> peer_dst_off;
    std::vector<long long> peer_ldW;                 // leading dimension of the peer's workspace (pushes into column k > 0)

    // CSR (local indices)
    int32_t* d_rowptr = nullptr;
    int32_t* d_colind = nullptr;
    double* d_val = nullptr;
    int csr_lanes = 8;

    // long rows (> kLongRow entries, power-law hubs): kept out of the main layout and processed by one CTA per segment of
    // kLongSeg entries (deterministic two-stage sum), so that no warp of the main kernel crawls through a 10^6-entry row
    int64_t n_long = 0, n_long_seg = 0;
    int32_t* d_long_row = nullptr;                   // local row index per long row
    int32_t* d_long_seg0 = nullptr;                  // first segment of long row r (n_long + 1 entries)
    int32_t* d_long_segptr = nullptr;                // first entry of segment g (n_long_seg + 1 entries)
    int32_t* d_long_segrow = nullptr;                // long-row slot of segment g
    int32_t* d_long_col = nullptr;
    double* d_long_val = nullptr;
    double* d_long_part = nullptr;                   // per-segment partial sums

    // SELL-32-sigma (column-major inside a slice of 32 rows)
    int64_t sell_slices = 0, sell_padded = 0;
    int sell_sigma = 32;
    int32_t* d_slice_ptr = nullptr;                  // entry offset of slice / 32 (i.e. first "row" of width)
    int32_t* d_sell_col = nullptr;
    double* d_sell_val = nullptr;
    int32_t* d_perm = nullptr;                       // sorted position -> local row (NULL: identity)

    // dictionary-coded SELL: 8 code bytes per lane per block, dictionary of {value, offset} pairs
    uint8_t* d_codes = nullptr;
    double* d_dict = nullptr;
    int dict_size = 0;
    double dict_uniform = 0.0;                        // fraction of (block, slot) positions whose 32 lanes share one code
    struct alignas(16) HostDictEnt { double v; int offb; int pad; };
    struct { HostDictEnt e[256]; } h_dict[1] = {};       // host copy, passed to the kernels as a constant-bank parameter
    // slice patterns: a slice whose 32 rows hold the SAME (offset, value) entries -- possibly on a subset of the lanes, e.g. the
    // x-boundary rows of a stencil miss one neighbour -- is described by one of <= 32 patterns of <= 8 {value, byte offset, lane
    // mask} entries instead of 256 code bytes: the kernel then neither reads codes nor decodes them (mpk.cu, k_spmv_selp)
    struct alignas(16) HostPatEnt { double v; int offb; unsigned mask; };
    // kernel parameter block: the entries of pattern 0 (the most frequent one: the interior of a stencil) and, for every pattern,
    // one lane mask per entry of pattern 0 -- a pattern must be a sub-pattern of pattern 0 (same offsets and values on fewer
    // lanes: boundary rows), otherwise its slices take the coded path
    struct { HostPatEnt e0[8]; unsigned mask[32][8]; } h_pat[1] = {};
    int n_pat = 0;
    int pat_cnt0 = 0;                                 // entries of pattern 0; 0: no usable pattern 0 (it must cover all 32 lanes)
    int pat_phase = 0;                                // slice pairing of k_spmv_selp: items are slices (2j - phase, 2j - phase + 1)
    double pat_cover = 0.0;                           // fraction of the slices that have a pattern
    uint8_t* d_slice_pat = nullptr;                   // per slice: pattern number, 255 = none (coded path)
    // ... and its TMA-staged kernel: x segments of a CTA's row block (merged over overlapping offsets)
    int xs_rows = 0;                                  // rows per CTA (0: staged kernel not applicable)
    int xs_groups = 0;
    int xs_omin[8] = {}, xs_len[8] = {}, xs_base[8] = {};
    int xs_total = 0;                                 // doubles of shared memory for the segments
    int* d_xs_off = nullptr;                          // per code: position of x[row+offset] relative to (row - r0)

    unsigned long long* d_gridbar = nullptr;           // grid-barrier counter of the fused MPK launch (monotonic)
    unsigned long long gridbar_base = 0;

    // basis workspace n_loc x (s_max+1), ghosts included
    double* d_W = nullptr;                            // d_W_alloc + W_pad: (d_W + own_off) is 16-byte aligned
    double* d_W_alloc = nullptr;
    int W_pad = 0;
    int64_t ldW = 0;
};

namespace calz {
int p2p_halo_setup(calz_mat* m);
void p2p_halo_teardown(calz_mat* m);
int p2p_halo_exchange(calz_mat* m, double* w, int col);
int p2p_halo_ack(calz_mat* m);
}  // namespace calz
