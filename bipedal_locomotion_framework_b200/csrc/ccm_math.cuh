// ccm_math.cuh -- per-contact closed form of ContinuousContactModel, register-resident FP64.
//
// Mathematics follows the reference (src/ContactModels/src/ContinuousContactModel.cpp):
//   wrench     :79-108    autonomous dynamics :110-146    control matrix :148-171
//   regressor  :223-254   point force/torque  :173-221
// but is derived independently in cross-product form: S(a)b = a x b, S(e)S(e)w = e x (e x w),
// Rdot.col(i) = w x e_i, S(e)^2 = e e^T - |e|^2 I.  One evaluation shares every sub-expression
// across the four outputs (the reference recomputes them per getter).  The reference's sign quirk
// is kept: wrench and regressor use |R22|, autonomous dynamics and control matrix use signed R22.
//
// Everything here is plain FP64 FMA-pipe work (DFMA/DMUL/DADD); tensor cores do not apply to
// per-contact 3-vector algebra.
#pragma once

#include <cuda_runtime.h>

namespace blfccm {

enum : unsigned { M_WRENCH = 1u, M_AUTODYN = 2u, M_CTRL = 4u, M_REGRESSOR = 8u };

// Parameters of one contact, with the products every output needs.
struct Prm {
    double L2;    // length^2
    double W2;    // width^2
    double A;     // length*width
    double A12;   // length*width/12
    double k;     // spring_coeff
    double b;     // damper_coeff
};

__host__ __device__ __forceinline__ Prm make_prm(double length, double width, double k, double b)
{
    Prm p;
    p.L2 = length * length;
    p.W2 = width * width;
    p.A = length * width;
    p.A12 = p.A / 12.0;
    p.k = k;
    p.b = b;
    return p;
}

struct V3 {
    double x, y, z;
};

__device__ __forceinline__ V3 cross(const V3& a, const V3& b)
{
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Cross product with every product rounded on its own (no FMA contraction): a x a is then exactly
// zero, as in the reference's non-fused arithmetic.  Used for S(e_i) R0.col(i), which cancels
// exactly when the foot sits at its null-force orientation (R == R0), a common resting state.
__device__ __forceinline__ V3 cross_exact(const V3& a, const V3& b)
{
    return V3{__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)),
              __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
              __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}
__device__ __forceinline__ V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, const V3& a) { return V3{s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 neg(const V3& a) { return V3{-a.x, -a.y, -a.z}; }

// Live inputs of one contact state (27 doubles when everything is requested).
struct State {
    V3 v, w;      // mixed twist: linear, angular
    V3 p, p0;     // link position, null-force position
    V3 e1, e2;    // R.col(0), R.col(1)
    double R02, R12, R22;  // R.col(2)   (R02, R12 only feed Rdot(2,2))
    V3 n1, n2;    // R0.col(0), R0.col(1)
};

// Results.  g is stored compactly: gd = the three equal top-left diagonal entries,
// gs = bottom-right symmetric 3x3 as {xx, xy, xz, yy, yz, zz}; every other entry of the dense
// 6x6 is a structural +0.0 (ContinuousContactModel.cpp:18,165-170).
struct Result {
    V3 force, torque;       // wrench
    V3 fhead, ftail;        // autonomous dynamics
    double gd;
    double gs[6];
    V3 y_fk, y_fb, y_tk, y_tb;  // regressor corners: (force|torque) x (spring|damper) columns
    V3 t1, t2;              // e1 x w, e2 x w = -Rdot.col(0), -Rdot.col(1): reused by the fused rollout
};

template <unsigned MASK>
__device__ __forceinline__ void eval_contact(const State& s, const Prm& q, Result& r)
{
    constexpr bool kW = (MASK & M_WRENCH) != 0;
    constexpr bool kA = (MASK & M_AUTODYN) != 0;
    constexpr bool kC = (MASK & M_CTRL) != 0;
    constexpr bool kR = (MASK & M_REGRESSOR) != 0;

    const double c = s.R22;
    const double ac = fabs(c);

    // shared by wrench / autodyn / regressor
    V3 d, sd, t1, t2, u1, u2, m1, m2, Tb;
    if constexpr (kW || kA || kR) {
        d = s.p0 - s.p;
        t1 = cross(s.e1, s.w);   // e1 x w   ( = -Rdot.col(0) )
        t2 = cross(s.e2, s.w);
        u1 = cross(s.e1, t1);    // S(e1)^2 w
        u2 = cross(s.e2, t2);
        m1 = cross_exact(s.e1, s.n1);  // S(e1) R0.col(0)
        m2 = cross_exact(s.e2, s.n2);
        r.t1 = t1;
        r.t2 = t2;
    }
    if constexpr (kW || kA) {
        sd = q.k * d - q.b * s.v;                       // k (p0 - p) - b v
        const V3 T1 = q.b * u1 + q.k * m1;
        const V3 T2 = q.b * u2 + q.k * m2;
        Tb = q.L2 * T1 + q.W2 * T2;                     // the torque bracket of :102-107
    }
    if constexpr (kW) {
        r.force = (ac * q.A) * sd;
        r.torque = (ac * q.A12) * Tb;
    }
    if constexpr (kA) {
        // Rdot(2,2) = (w x R.col(2)).z
        const double cdot = s.w.x * s.R12 - s.w.y * s.R02;
        r.fhead = q.A * (cdot * sd - (c * q.k) * s.v);
        const V3 d1 = neg(t1), d2 = neg(t2);            // Rdot.col(i) = w x e_i
        // k S(de) n + b (S(de)S(e) + S(e)S(de)) w
        const V3 Q1 = q.k * cross(d1, s.n1) + q.b * (cross(d1, t1) + cross(s.e1, cross(d1, s.w)));
        const V3 Q2 = q.k * cross(d2, s.n2) + q.b * (cross(d2, t2) + cross(s.e2, cross(d2, s.w)));
        const V3 Q = q.L2 * Q1 + q.W2 * Q2;
        r.ftail = q.A12 * (cdot * Tb + c * Q);
    }
    if constexpr (kC || kR) {
        // M = L^2 S(e1)^2 + W^2 S(e2)^2 (symmetric)
        const V3 &a = s.e1, &e = s.e2;
        const double mxx = -(q.L2 * (a.y * a.y + a.z * a.z) + q.W2 * (e.y * e.y + e.z * e.z));
        const double myy = -(q.L2 * (a.x * a.x + a.z * a.z) + q.W2 * (e.x * e.x + e.z * e.z));
        const double mzz = -(q.L2 * (a.x * a.x + a.y * a.y) + q.W2 * (e.x * e.x + e.y * e.y));
        const double mxy = q.L2 * (a.x * a.y) + q.W2 * (e.x * e.y);
        const double mxz = q.L2 * (a.x * a.z) + q.W2 * (e.x * e.z);
        const double myz = q.L2 * (a.y * a.z) + q.W2 * (e.y * e.z);
        if constexpr (kC) {
            r.gd = -(q.A * q.b) * c;
            const double sc = q.A12 * c * q.b;
            r.gs[0] = sc * mxx; r.gs[1] = sc * mxy; r.gs[2] = sc * mxz;
            r.gs[3] = sc * myy; r.gs[4] = sc * myz; r.gs[5] = sc * mzz;
        }
        if constexpr (kR) {
            const double sf = ac * q.A, st = q.A12 * ac;
            r.y_fk = sf * d;
            r.y_fb = (-sf) * s.v;
            r.y_tk = st * (q.L2 * m1 + q.W2 * m2);
            r.y_tb = st * V3{mxx * s.w.x + mxy * s.w.y + mxz * s.w.z,
                             mxy * s.w.x + myy * s.w.y + myz * s.w.z,
                             mxz * s.w.x + myz * s.w.y + mzz * s.w.z};
        }
    }
}

// Which of the 30 SoA planes can affect the requested outputs (bit i = plane i).
__host__ __device__ constexpr unsigned live_planes(unsigned mask)
{
    // v w p | R col0,col1 (9,10,12,13,15,16) | R22 (17) | p0 | R0 col0,col1 (21,22,24,25,27,28)
    unsigned all = 0;
    const unsigned vwp = 0x1FFu;                       // planes 0..8
    const unsigned e12 = (1u << 9) | (1u << 10) | (1u << 12) | (1u << 13) | (1u << 15) | (1u << 16);
    const unsigned r22 = 1u << 17;
    const unsigned r0212 = (1u << 11) | (1u << 14);
    const unsigned p0 = (1u << 18) | (1u << 19) | (1u << 20);
    const unsigned n12 = (1u << 21) | (1u << 22) | (1u << 24) | (1u << 25) | (1u << 27) | (1u << 28);
    if (mask & (M_WRENCH | M_AUTODYN | M_REGRESSOR)) all |= vwp | e12 | r22 | p0 | n12;
    if (mask & M_AUTODYN) all |= r0212;
    if (mask & M_CTRL) all |= e12 | r22;
    return all;
}

}  // namespace blfccm
