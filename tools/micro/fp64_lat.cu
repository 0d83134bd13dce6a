// Dependent-issue latency of the FP64 pipe and of an IEEE double division / reciprocal on this GPU
// (one warp, one block): cycles per dependent instruction.  nvcc -arch=sm_100a -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(double* out, long long* cycles, double a, double b, int iters)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) x = fma(x, y, a);                 // DFMA chain
            else if (MODE == 1) x = y / x + a;               // IEEE division + DADD
            else if (MODE == 2) x = __drcp_rn(x) + a;        // IEEE reciprocal + DADD
            else if (MODE == 3) {                            // approximate reciprocal: MUFU.RCP64H + 2 Newton steps
                double r;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
                double e = fma(-x, r, 1.0);
                r = fma(r, e, r);
                e = fma(-x, r, 1.0);
                r = fma(r, e, r);
                x = r + a;
            } else if (MODE == 4) x = x * y;                  // DMUL chain
            else if (MODE == 5) x = x + y;                    // DADD chain
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, int deps_per_op)
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 32 * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    chain<MODE><<<1, 32>>>(out, cyc, 1.25, 0.75, iters);
    chain<MODE><<<1, 32>>>(out, cyc, 1.25, 0.75, iters);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    printf("%-52s %8.1f cycles per op (%d dependent ops in it)\n", name, double(c) / (iters * 16.0), deps_per_op);
    cudaFree(out);
    cudaFree(cyc);
}

// Issue throughput: W warps in one block (one SM), each with 8 independent DFMA chains.
// cycles per warp-DFMA per scheduler tells how wide the FP64 pipe of one SM sub-partition is.
__global__ void tput(double* out, long long* cycles, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x * 1e-9 + j;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, a);
    }
    __syncthreads();
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

void run_tput(int warps)
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 32 * warps * sizeof(double));
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    tput<<<1, 32 * warps>>>(out, cyc, 1.25, 0.75, iters);
    tput<<<1, 32 * warps>>>(out, cyc, 1.25, 0.75, iters);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_warp = double(c) / (iters * 32.0);
    printf("DFMA throughput, %2d warps on one SM, 8 independent chains each: %6.2f cycles per DFMA per warp, %6.2f warp-DFMA per cycle per SM\n",
           warps, per_warp, warps / per_warp);
    cudaFree(out);
    cudaFree(cyc);
}

// Which warps of a CTA share a scheduler (SM sub-partition)?  Only the warps named in `mask` run the
// independent-DFMA loop; two warps on the same sub-partition halve each other's rate.
__global__ void tput_mask(long long* cycles, unsigned* smid, double* out, unsigned mask, double a, double b, int iters)
{
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        unsigned s;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
        smid[blockIdx.x] = s;
    }
    if (!((mask >> warp) & 1u)) return;
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x * 1e-9 + j;
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, a);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
}

void run_mask(int blocks, int warps, unsigned mask, const char* what)
{
    long long* cyc;
    unsigned* smid;
    double* out;
    cudaMalloc(&cyc, blocks * 32 * sizeof(long long));
    cudaMalloc(&smid, blocks * sizeof(unsigned));
    cudaMalloc(&out, size_t(blocks) * warps * 32 * sizeof(double));
    const int iters = 4000;
    for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(cyc, 0, blocks * 32 * sizeof(long long));
        tput_mask<<<blocks, 32 * warps>>>(cyc, smid, out, mask, 1.25, 0.75, iters);
    }
    static long long hc[1024 * 32];
    static unsigned hs[1024];
    cudaMemcpy(hc, cyc, blocks * 32 * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(hs, smid, blocks * sizeof(unsigned), cudaMemcpyDeviceToHost);
    // blocks that had an SM to themselves vs blocks that shared one
    int per_sm[256] = {0};
    for (int b = 0; b < blocks; ++b) per_sm[hs[b] & 255]++;
    double alone = 0, shared = 0;
    int na = 0, ns = 0;
    for (int b = 0; b < blocks; ++b)
        for (int w = 0; w < warps; ++w)
            if ((mask >> w) & 1u) {
                const double c = double(hc[b * 32 + w]) / (iters * 32.0);
                if (per_sm[hs[b] & 255] > 1) shared += c, ns++;
                else alone += c, na++;
            }
    printf("%-58s", what);
    if (na) printf("  %5.2f cycles/DFMA (CTA alone on its SM)", alone / na);
    if (ns) printf("  %5.2f cycles/DFMA (CTAs sharing an SM, %d warps)", shared / ns, ns);
    printf("\n");
    cudaFree(cyc);
    cudaFree(smid);
    cudaFree(out);
}

int main()
{
    run_mask(1, 16, 0x0001, "1 CTA, warp 0 only");
    run_mask(1, 16, 0x0003, "1 CTA, warps 0 1");
    run_mask(1, 16, 0x0005, "1 CTA, warps 0 2");
    run_mask(1, 16, 0x0011, "1 CTA, warps 0 4");
    run_mask(1, 16, 0x0021, "1 CTA, warps 0 5");
    run_mask(1, 16, 0x0101, "1 CTA, warps 0 8");
    run_mask(1, 16, 0x1111, "1 CTA, warps 0 4 8 12");
    run_mask(1, 16, 0x000f, "1 CTA, warps 0 1 2 3");
    run_mask(296, 4, 0x1, "296 CTAs of 4 warps, warp 0 of each");
    run_mask(296, 5, 0x1, "296 CTAs of 5 warps, warp 0 of each");
    run_mask(296, 4, 0xf, "296 CTAs of 4 warps, all warps");
    run_mask(256, 4, 0x1, "256 CTAs of 4 warps, warp 0 of each");
    run_mask(256, 8, 0x3, "256 CTAs of 8 warps, warps 0 1");
    for (int w : {1, 2, 4, 8, 16}) run_tput(w);
    for (int w : {1, 2, 4, 8, 16}) run_tput(w);
    run<0>("DFMA dependent chain", 1);
    run<4>("DMUL dependent chain", 1);
    run<5>("DADD dependent chain", 1);
    run<1>("IEEE division y/x + DADD", 2);
    run<2>("IEEE reciprocal __drcp_rn + DADD", 2);
    run<3>("rcp.approx.ftz.f64 + 2 Newton steps + DADD", 6);
    return 0;
}
