// Device-side helpers of the peer-memory collectives (shared by p2p.cu and the fused reduce + all-reduce of tiles.cu).
#pragma once
#include "calz_internal.h"

namespace calz {

constexpr unsigned long long kSpinLimit = 100ull * 1000ull * 1000ull;     // ~1 min; a healthy wait is microseconds

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long seq, int* err) {
    unsigned long long it = 0;
    while (ld_acquire_sys(flag) < seq) {
        if (++it > kSpinLimit || (it % 4096 == 0 && *(volatile int*)err)) {
            *(volatile int*)err = 1;
            return false;
        }
    }
    return true;
}

struct PeerPtrs {
    double* mbox[kMaxPeers];
    unsigned long long* flags[kMaxPeers];
};

// everything a kernel needs to take part in mailbox all-reduce number `seq` (P == 1: no communicator / not applicable)
struct ArArgs {
    int P, me;
    unsigned long long seq;
    PeerPtrs peers;
    double* my_mbox;
    unsigned long long* my_flags;
    int* err;
    unsigned int* ticket;
};

// fills `a` for the next mailbox all-reduce of the context and advances its sequence number (p2p.cu)
void p2p_next_allreduce(calz_ctx* ctx, ArArgs* a);

}  // namespace calz
