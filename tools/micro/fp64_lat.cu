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

int main()
{
    run<0>("DFMA dependent chain", 1);
    run<4>("DMUL dependent chain", 1);
    run<5>("DADD dependent chain", 1);
    run<1>("IEEE division y/x + DADD", 2);
    run<2>("IEEE reciprocal __drcp_rn + DADD", 2);
    run<3>("rcp.approx.ftz.f64 + 2 Newton steps + DADD", 6);
    return 0;
}
