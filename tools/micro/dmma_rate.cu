// dmma_rate.cu -- issue rate and dependent latency of mma.sync.m8n8k4.f64 (DMMA) on one SM and on the whole
// GPU, beside DFMA: input for a blocked mass-matrix factorisation whose trailing updates would take their
// operands from registers across the warp instead of shared-memory broadcasts (development aid).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate tools/micro/dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS, bool USE_MMA>
__global__ void rate_kernel(double* out, int iters, long long* cycles)
{
    double c[CHAINS][2];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) c[q][0] = c[q][1] = threadIdx.x + q;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < CHAINS; ++q) {
            if (USE_MMA) dmma(c[q][0], c[q][1], a, b);
            else {
                c[q][0] = fma(a, b, c[q][0]);
                c[q][1] = fma(a, b, c[q][1]);
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) s += c[q][0] + c[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS, bool USE_MMA>
void run(const char* what, int blocks, int threads)
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaMalloc(&cyc, 8);
    const int iters = 20000;
    rate_kernel<CHAINS, USE_MMA><<<blocks, threads>>>(out, 100, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<CHAINS, USE_MMA><<<blocks, threads>>>(out, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = double(iters) * CHAINS;                  // warp-level ops per warp
    const double warps = double(blocks) * threads / 32;
    // an m8n8k4 DMMA = 256 multiply-adds per warp instruction; the DFMA pair here = 64
    const double fma_per_op = USE_MMA ? 256.0 : 64.0;
    printf("%-44s %2d chains  %6.2f cycles/op/warp  %8.2f TFLOP/s (whole launch)\n", what, CHAINS, c / ops,
           2.0 * fma_per_op * ops * warps / (ms * 1e-3) / 1e12);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    run<1, true>("DMMA m8n8k4, 1 warp, dependent", 1, 32);
    run<8, true>("DMMA m8n8k4, 1 warp, 8 independent", 1, 32);
    run<8, true>("DMMA m8n8k4, 4 warps on one SM", 1, 128);
    run<8, true>("DMMA m8n8k4, 16 warps on one SM", 1, 512);
    run<8, true>("DMMA m8n8k4, 16 warps on every SM", sms, 512);
    run<1, false>("DFMA pair, 1 warp, dependent", 1, 32);
    run<8, false>("DFMA pairs, 16 warps on one SM", 1, 512);
    run<8, false>("DFMA pairs, 16 warps on every SM", sms, 512);
    return 0;
}
