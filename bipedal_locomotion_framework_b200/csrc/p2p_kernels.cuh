// p2p_kernels.cuh -- arg-min exchange over NVLink / NVSwitch peer memory (one process per GPU).
//
// The only cross-GPU step of the path is the global arg-min of every rank's 16-byte (cost, index)
// pair.  With NCCL that is an all-gather whose cost is pure launch/protocol latency (+13 us per
// step measured on 2 B200s, +22 us on 8).  Here every rank owns a small MAILBOX in its own HBM,
// mapped into every peer through CUDA IPC; ONE single-warp kernel per rank
//   1. stores its pair into slot [parity][rank] of every peer's mailbox (remote stores over
//      NVLink) as four 8-byte words, each = 4 data bytes + the 32-bit epoch ("LL" framing: an
//      8-byte store is atomic, so data and flag arrive together and no system-wide fence -- a
//      full NVLink round trip -- is needed between them),
//   2. spins on the 4 x nranks words of its OWN mailbox until they all carry this epoch,
//   3. reduces the nranks pairs with the lowest-index tie-break and writes the global best.
// No host round trip, no collective library.  Slots are double-buffered on the epoch parity: a rank
// can be at most one exchange ahead of a peer (it needs the peer's flag of exchange e to finish e),
// so the pair of exchange e is never overwritten before every peer has read it.
// A peer that never shows up would hang the spin: it is bounded (~2 s) and then reports index -2.
// Included by ccm_kernels.cuh right after CostIdx / better() / warp_best() (it needs them), so that
// the cost-reduction kernel can run the exchange in its last block: reduce + exchange in ONE launch.
#pragma once

namespace blfccm {

constexpr int kP2pMaxRanks = 32;

struct alignas(32) P2pSlot {
    unsigned long long w[4];   // w[k] = (epoch32 << 32) | k-th 32-bit quarter of {cost, idx}
};

struct P2pArgs {
    P2pSlot* peer[kP2pMaxRanks];   // every rank's mailbox as mapped in THIS process ([2][nranks])
    const CostIdx* mine;           // this rank's pair
    CostIdx* out;                  // global best
    unsigned long long epoch;      // >= 1, same on every rank
    int nranks;
    int rank;
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One warp: publish `me`, wait for every rank's pair of this epoch, return the global best
// (index -1: nothing to compare anywhere; -2: a peer did not arrive in time).
__device__ __forceinline__ CostIdx p2p_exchange_warp(const P2pArgs& a, const CostIdx me, int lane)
{
    const int parity = static_cast<int>(a.epoch & 1ull);
    const unsigned long long tag = (a.epoch & 0xffffffffull) << 32;
    // 1. publish to every peer (lane r -> rank r's mailbox)
    if (lane < a.nranks) {
        P2pSlot* s = a.peer[lane] + parity * a.nranks + a.rank;
        const unsigned long long c = static_cast<unsigned long long>(__double_as_longlong(me.cost));
        const unsigned long long i = static_cast<unsigned long long>(me.idx);
        st_sys_u64(&s->w[0], tag | (c & 0xffffffffull));
        st_sys_u64(&s->w[1], tag | (c >> 32));
        st_sys_u64(&s->w[2], tag | (i & 0xffffffffull));
        st_sys_u64(&s->w[3], tag | (i >> 32));
    }
    // 2. wait for every rank's pair of this epoch in the local mailbox
    CostIdx b{__longlong_as_double(0x7ff0000000000000LL), 0x7fffffffffffffffLL};
    bool timeout = false;
    if (lane < a.nranks) {
        const P2pSlot* s = a.peer[a.rank] + parity * a.nranks + lane;
        const long long t0 = clock64();
        unsigned long long w0, w1, w2, w3;
        for (;;) {
            w0 = ld_sys_u64(&s->w[0]);
            w1 = ld_sys_u64(&s->w[1]);
            w2 = ld_sys_u64(&s->w[2]);
            w3 = ld_sys_u64(&s->w[3]);
            const unsigned long long hi = 0xffffffff00000000ull;
            if ((w0 & hi) == tag && (w1 & hi) == tag && (w2 & hi) == tag && (w3 & hi) == tag) break;
            if (clock64() - t0 > 4000000000LL) {
                timeout = true;
                break;
            }
        }
        if (!timeout) {
            b.cost = __longlong_as_double(static_cast<long long>((w0 & 0xffffffffull) | (w1 << 32)));
            b.idx = static_cast<long long>((w2 & 0xffffffffull) | (w3 << 32));
            if (b.idx < 0) {   // a rank with nothing to compare
                b.cost = __longlong_as_double(0x7ff0000000000000LL);
                b.idx = 0x7fffffffffffffffLL;
            }
        }
    }
    const bool any_timeout = __any_sync(0xffffffffu, timeout);
    // 3. combine
    b = warp_best(b);
    if (any_timeout) {
        b.cost = __longlong_as_double(0x7ff8000000000000LL);
        b.idx = -2;
    } else if (b.idx == 0x7fffffffffffffffLL) {
        b.idx = -1;
    }
    return b;
}

__global__ void __launch_bounds__(32)
ccm_p2p_exchange_kernel(const __grid_constant__ P2pArgs a)
{
    const int lane = threadIdx.x;
    const CostIdx b = p2p_exchange_warp(a, *a.mine, lane);
    if (lane == 0) *a.out = b;
}

}  // namespace blfccm
