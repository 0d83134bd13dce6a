// Cycles per ForwardEuler step of the pose recurrence for ONE warp alone on an SM sub-partition
// (what bounds the producer warp of the warp-specialised rollout kernels), for several ways of
// writing the same step.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bipedal_locomotion_framework_b200/csrc
//                               -o kin_step_lat tools/micro/kin_step_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "ccm_math.cuh"
namespace blfccm {
constexpr int kWarp = 32;
struct CostIdx { double cost; long long idx; };
}
using namespace blfccm;

struct Pose { V3 p, c0, c1, c2; };

__device__ __forceinline__ double rcp2(double x)   // seed + two Newton steps (the shipped fast_rcp)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double dot3(const V3& a, const V3& b) { return fma(a.z, b.z, fma(a.y, b.y, __dmul_rn(a.x, b.x))); }

// V0: the shipped step
__device__ __forceinline__ void step_v0(Pose& s, const V3& v, const V3& w, double alpha, double dthr, double dT)
{
    s.p = V3{fma(dT, v.x, s.p.x), fma(dT, v.y, s.p.y), fma(dT, v.z, s.p.z)};
    const V3 x0 = cross(w, s.c0), x1 = cross(w, s.c1), x2 = cross(w, s.c2);
    const V3 k0 = cross(s.c1, s.c2), k1 = cross(s.c2, s.c0), k2 = cross(s.c0, s.c1);
    const double beta = __dmul_rn(dthr, rcp2(dot3(s.c0, k0)));
    auto col = [&](const V3& c, const V3& x, const V3& k) {
        return V3{fma(beta, k.x, fma(dT, x.x, __dmul_rn(alpha, c.x))), fma(beta, k.y, fma(dT, x.y, __dmul_rn(alpha, c.y))),
                  fma(beta, k.z, fma(dT, x.z, __dmul_rn(alpha, c.z)))};
    };
    const V3 n0 = col(s.c0, x0, k0), n1 = col(s.c1, x1, k1), n2 = col(s.c2, x2, k2);
    s.c0 = n0; s.c1 = n1; s.c2 = n2;
}

// V1: dT w formed once, the cross product folded into the update (4 instead of 5 operations per component)
__device__ __forceinline__ V3 col_v1(const V3& c, const V3& dw, const V3& k, double alpha, double beta)
{
    return V3{fma(beta, k.x, fma(dw.y, c.z, fma(-dw.z, c.y, __dmul_rn(alpha, c.x)))),
              fma(beta, k.y, fma(dw.z, c.x, fma(-dw.x, c.z, __dmul_rn(alpha, c.y)))),
              fma(beta, k.z, fma(dw.x, c.y, fma(-dw.y, c.x, __dmul_rn(alpha, c.z))))};
}
template <int RCP>
__device__ __forceinline__ double beta_of(double det, double dthr)
{
    if (RCP == 0) return __dmul_rn(dthr, rcp2(det));
    // cubic step: r = r0 (1 + e + e^2), e = 1 - det r0; |e| <= 2^-20 -> error e^3; dthr folded in
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(det));
    const double e = fma(-det, r0, 1.0);
    const double b0 = __dmul_rn(dthr, r0);
    const double t = fma(e, e, e);
    return fma(b0, t, b0);
}
template <int RCP>
__device__ __forceinline__ void step_v1(Pose& s, const V3& v, const V3& w, double alpha, double dthr, double dT)
{
    s.p = V3{fma(dT, v.x, s.p.x), fma(dT, v.y, s.p.y), fma(dT, v.z, s.p.z)};
    const V3 dw{__dmul_rn(dT, w.x), __dmul_rn(dT, w.y), __dmul_rn(dT, w.z)};
    const V3 k0 = cross(s.c1, s.c2);
    const double beta = beta_of<RCP>(dot3(s.c0, k0), dthr);
    const V3 k1 = cross(s.c2, s.c0), k2 = cross(s.c0, s.c1);
    const V3 n0 = col_v1(s.c0, dw, k0, alpha, beta), n1 = col_v1(s.c1, dw, k1, alpha, beta), n2 = col_v1(s.c2, dw, k2, alpha, beta);
    s.c0 = n0; s.c1 = n1; s.c2 = n2;
}
// no Baumgarte term
__device__ __forceinline__ void step_n0(Pose& s, const V3& v, const V3& w, double dT)
{
    s.p = V3{fma(dT, v.x, s.p.x), fma(dT, v.y, s.p.y), fma(dT, v.z, s.p.z)};
    const V3 x0 = cross(w, s.c0), x1 = cross(w, s.c1), x2 = cross(w, s.c2);
    s.c0 = V3{fma(dT, x0.x, s.c0.x), fma(dT, x0.y, s.c0.y), fma(dT, x0.z, s.c0.z)};
    s.c1 = V3{fma(dT, x1.x, s.c1.x), fma(dT, x1.y, s.c1.y), fma(dT, x1.z, s.c1.z)};
    s.c2 = V3{fma(dT, x2.x, s.c2.x), fma(dT, x2.y, s.c2.y), fma(dT, x2.z, s.c2.z)};
}
__device__ __forceinline__ void step_n1(Pose& s, const V3& v, const V3& w, double dT)
{
    s.p = V3{fma(dT, v.x, s.p.x), fma(dT, v.y, s.p.y), fma(dT, v.z, s.p.z)};
    const V3 dw{__dmul_rn(dT, w.x), __dmul_rn(dT, w.y), __dmul_rn(dT, w.z)};
    auto col = [&](const V3& c) {
        return V3{fma(dw.y, c.z, fma(-dw.z, c.y, c.x)), fma(dw.z, c.x, fma(-dw.x, c.z, c.y)), fma(dw.x, c.y, fma(-dw.y, c.x, c.z))};
    };
    const V3 n0 = col(s.c0), n1 = col(s.c1), n2 = col(s.c2);
    s.c0 = n0; s.c1 = n1; s.c2 = n2;
}

template <int V, int NCH, bool STORE>
__global__ void __launch_bounds__(32) run(double* out, long long* cycles, const double* tw_g, int boxes, double dT, double half_rho)
{
    __shared__ double tw[6 * 8 * 32];
    __shared__ double2 stage[8 * 5 * 32];
    const int lane = threadIdx.x;
    for (int i = lane; i < 6 * 8 * 32; i += 32) tw[i] = tw_g[i];
    __syncwarp();
    Pose s[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        s[j].p = V3{0.1 * lane, 0.2, 0.3 + j};
        s[j].c0 = V3{1.0, 1e-3 * lane, 0.0};
        s[j].c1 = V3{-1e-3 * lane, 1.0, 1e-4 * j};
        s[j].c2 = V3{0.0, -1e-4 * j, 1.0};
    }
    const double alpha = 1.0 - dT * half_rho, dthr = dT * half_rho;
    const long long t0 = clock64();
#pragma unroll 1
    for (int b = 0; b < boxes; ++b) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double* r = tw + q * 32 + lane;
            const V3 v{r[0], r[256], r[512]};
            const V3 w{r[768], r[1024], r[1280]};
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                if (STORE && j == 0) {
                    double2* o = stage + q * 160 + lane;
                    o[0] = make_double2(s[j].p.x, s[j].p.y);
                    o[32] = make_double2(s[j].p.z, s[j].c0.x);
                    o[64] = make_double2(s[j].c0.y, s[j].c0.z);
                    o[96] = make_double2(s[j].c1.x, s[j].c1.y);
                    o[128] = make_double2(s[j].c1.z, s[j].c2.z);
                }
                if (V == 0) step_v0(s[j], v, w, alpha, dthr, dT);
                else if (V == 1) step_v1<0>(s[j], v, w, alpha, dthr, dT);
                else if (V == 2) step_v1<1>(s[j], v, w, alpha, dthr, dT);
                else if (V == 10) step_n0(s[j], v, w, dT);
                else if (V == 11) step_n1(s[j], v, w, dT);
            }
        }
    }
    const long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
        acc += s[j].p.x + s[j].c0.x + s[j].c0.y + s[j].c0.z + s[j].c1.x + s[j].c1.y + s[j].c1.z + s[j].c2.x + s[j].c2.y + s[j].c2.z;
    out[blockIdx.x * 32 + lane] = acc + (STORE ? stage[lane].x : 0.0);
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V, int NCH, bool STORE>
void go(const char* name, const double* tw)
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 32 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    const int boxes = 250;
    for (int rep = 0; rep < 2; ++rep) run<V, NCH, STORE><<<1, 32>>>(out, cyc, tw, boxes, 0.01, 0.005);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double o = 0;
    cudaMemcpy(&o, out, sizeof(o), cudaMemcpyDeviceToHost);
    printf("%-78s %7.1f cycles per step per chain-set, %7.1f per chain   (check %.17g)\n", name, double(c) / (boxes * 8.0),
           double(c) / (boxes * 8.0 * NCH), o);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    double h[6 * 8 * 32];
    for (int i = 0; i < 6 * 8 * 32; ++i) h[i] = 0.3 * ((i * 2654435761u >> 8) % 1000) / 1000.0 - 0.15;
    double* tw;
    cudaMalloc(&tw, sizeof(h));
    cudaMemcpy(tw, h, sizeof(h), cudaMemcpyHostToDevice);
    go<0, 1, true>("Baumgarte, shipped step (75 FP64), pose stores", tw);
    go<0, 1, false>("Baumgarte, shipped step, no stores", tw);
    go<1, 1, true>("Baumgarte, dT w folded (69 FP64), pose stores", tw);
    go<2, 1, true>("Baumgarte, dT w folded + cubic reciprocal step with beta folded, pose stores", tw);
    go<2, 1, false>("Baumgarte, dT w folded + cubic reciprocal, no stores", tw);
    go<0, 2, true>("Baumgarte, shipped step, TWO chains per lane, pose stores for one", tw);
    go<2, 2, true>("Baumgarte, folded + cubic, TWO chains per lane", tw);
    go<2, 3, false>("Baumgarte, folded + cubic, THREE chains per lane", tw);
    go<10, 1, true>("no Baumgarte, shipped step (30 FP64), pose stores", tw);
    go<11, 1, true>("no Baumgarte, dT w folded (24 FP64), pose stores", tw);
    go<11, 2, true>("no Baumgarte, dT w folded, TWO chains per lane", tw);
    return 0;
}
