// Where do the warps of co-resident CTAs land?  Every warp records (%smid, %warpid) and stays alive
// for a while so that the second CTA of an SM arrives while the first one's warps still hold their
// slots.  warpid mod 4 is the SM sub-partition (scheduler).  nvcc -arch=sm_100a -o warp_slots warp_slots.cu
#include <cstdio>
#include <vector>
#include <map>
#include <cuda_runtime.h>

__global__ void probe(unsigned* smid, unsigned* warpid, long long spin, int smem_dummy)
{
    extern __shared__ double pad[];
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        unsigned s, w;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
        smid[blockIdx.x * nw + warp] = s;
        warpid[blockIdx.x * nw + warp] = w;
    }
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if (smem_dummy < 0) pad[threadIdx.x] = 1.0;
}

void run(int blocks, int warps, size_t smem)
{
    unsigned *s, *w;
    cudaMalloc(&s, blocks * warps * 4);
    cudaMalloc(&w, blocks * warps * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<blocks, 32 * warps, smem>>>(s, w, 200000, 0);
    cudaDeviceSynchronize();
    std::vector<unsigned> hs(blocks * warps), hw(blocks * warps);
    cudaMemcpy(hs.data(), s, hs.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost);
    std::map<unsigned, std::vector<int>> by_sm;
    for (int b = 0; b < blocks; ++b) by_sm[hs[b * warps]].push_back(b);
    int same = 0, diff = 0, pairs = 0;
    printf("%d CTAs x %d warps, %zu B dynamic smem:\n", blocks, warps, smem);
    int shown = 0;
    for (auto& kv : by_sm) {
        if (kv.second.size() < 2) continue;
        pairs++;
        const int a = kv.second[0], b = kv.second[1];
        if ((hw[a * warps] & 3) == (hw[b * warps] & 3)) same++; else diff++;
        if (shown++ < 4) {
            printf("  SM %3u: CTA %3d warpids", kv.first, a);
            for (int i = 0; i < warps; ++i) printf(" %u", hw[a * warps + i]);
            printf(" | CTA %3d warpids", b);
            for (int i = 0; i < warps; ++i) printf(" %u", hw[b * warps + i]);
            printf("\n");
        }
    }
    printf("  SMs with two CTAs: %d; warp 0 of both on the same sub-partition: %d, on different ones: %d\n", pairs, same, diff);
    cudaFree(s);
    cudaFree(w);
}

int main()
{
    run(256, 4, 50000);
    run(256, 5, 111000);
    run(256, 8, 111000);
    run(296, 5, 111000);
    return 0;
}
