// ccm_ptx.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, 1-D bulk async copies (TMA
// engine; SASS: UBLKCP), proxy fences.  All shared-memory operands are 32-bit shared-window
// addresses obtained with __cvta_generic_to_shared.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace blfccm {
namespace ptx {

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// make the init visible to the async proxy before the first bulk copy signals the barrier
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// plain arrive (release at CTA scope): the arriving thread's earlier shared-memory writes are
// visible to a thread that observes the phase completion with mbar_wait (acquire)
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // no PTX labels: the function is inlined several times into one kernel
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// Non-blocking phase test (acquire): true when the phase with this parity has completed.  Issued
// ahead of need, its latency (~90 cycles, like a try_wait that succeeds) overlaps other work; when
// it returns false the caller falls back to mbar_wait.
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}

// global -> shared::cta bulk copy, completion counted in bytes on an mbarrier.
// dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// shared::cta -> global bulk copy, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void bulk_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// all committed groups have finished READING shared memory (buffers reusable)
__device__ __forceinline__ void bulk_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// all committed groups complete (writes performed)
__device__ __forceinline__ void bulk_wait_all()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// order this thread's generic-proxy shared-memory writes before subsequent async-proxy reads
__device__ __forceinline__ void fence_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// launch_dependents: the next kernel of the stream (if launched with the programmatic-serialization
// attribute) may start occupying SM resources as they free up; it still blocks in grid_dep_wait()
// until THIS grid has completed and its memory is visible, so stream semantics are unchanged.
__device__ __forceinline__ void grid_dep_launch_dependents()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void grid_dep_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- per-thread async copies (LDGSTS): 8 bytes global -> shared, tracked by commit groups ------
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// at most N of this thread's committed groups still pending
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace ptx
}  // namespace blfccm
