// llt_bench.cu -- A/B timing of the warp-level mass-matrix solve variants (development aid).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DBLF_LLT_VAR=<k> and whatever switch the
//        kernel file currently reads] -DBLF_LLT_BENCH_CLASSES -I bipedal_locomotion_framework_b200/csrc
//        -o llt_bench_v<k> tools/micro/llt_bench.cu
// Includes the kernel translation unit itself, so what is timed is the product code.
#include "dyn_kernels.cu"

#ifndef BLF_LLT_VAR
#define BLF_LLT_VAR 0
#endif

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace blfccm;

__global__ void fill_kernel(double* M, double* known, long long n, int nc)
{
    const long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (s >= n) return;
    unsigned long long x = 0x9E3779B97F4A7C15ull * (s + 1);
    auto rnd = [&]() {
        x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
        return static_cast<double>((x * 0x2545F4914F6CDD1Dull) >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    };
    double* m = M + s * nc * nc;
    for (int i = 0; i < nc; ++i)
        for (int k = 0; k <= i; ++k) {
            const double v = (i == k) ? nc * 0.5 + rnd() : rnd();   // diagonally dominant: positive definite
            m[i * nc + k] = v;
            m[k * nc + i] = v;
        }
    for (int i = 0; i < nc; ++i) known[s * nc + i] = rnd() * 20.0;
}

int main(int argc, char** argv)
{
#if defined(BLF_LLT_ONLY_MID)
    const int sizes[][2] = {{16, 1 << 20}, {18, 1 << 20}, {19, 1 << 20}, {21, 1 << 20}, {23, 1 << 20}, {25, 1 << 20}, {27, 1 << 20}};
#elif defined(BLF_LLT_ONLY_WIDE)
    const int sizes[][2] = {{63, 1 << 17}, {56, 1 << 17}};
#elif defined(BLF_LLT_ONLY_29) || defined(BLF_LLT_ONLY_29_H16) || defined(BLF_LLT_ONLY_29_H32)
    const int sizes[][2] = {{29, 1 << 20}, {29, 409600}};
#else
    const int sizes[][2] = {{6, 1 << 22}, {12, 1 << 21}, {18, 1 << 20}, {23, 1 << 20}, {24, 1 << 20}, {29, 1 << 20}};
#endif
    for (auto& sz : sizes) {
        const int nc = sz[0];
        const long long n = sz[1];
        const int nb = 3;
        std::vector<double*> M(nb), out(nb);
        double* known;
        cudaMalloc(&known, n * nc * 8);
        for (int b = 0; b < nb; ++b) {
            cudaMalloc(&M[b], n * nc * nc * 8);
            cudaMalloc(&out[b], n * nc * 8);
            fill_kernel<<<(n + 127) / 128, 128>>>(M[b], known, n, nc);
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("fill failed\n"); return 1; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        auto run = [&](int b) {
            LltArgs a{M[b], nullptr, known, nullptr, out[b], n, nc};
            int path = 0, ncmax = 0;
            cudaError_t e = llt_solve_launch(a, 0, false, 0, &path, &ncmax);
            if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(1); }
        };
        for (int i = 0; i < 5; ++i) run(i % nb);
        cudaEventRecord(e0);
        const int iters = 20;
        for (int i = 0; i < iters; ++i) run(i % nb);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("run failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        double x0 = 0;
        cudaMemcpy(&x0, out[0], 8, cudaMemcpyDeviceToHost);
        const double tri = 8.0 * (nc * (nc + 1) / 2 + 2 * nc);
        printf("VAR %d nc %2d n %8lld  %9.1f us  %8.1f M systems/s  %6.1f GB/s lower-triangle  (x0 %.6g)\n", BLF_LLT_VAR, nc, n,
               ms * 1e3, n / ms / 1e3, n * tri / ms / 1e6, x0);
        for (int b = 0; b < nb; ++b) { cudaFree(M[b]); cudaFree(out[b]); }
        cudaFree(known);
    }
    return 0;
}
