// dyn_kernels.cu -- batched mass-matrix solve: the last step of
// System::FloatingBaseDynamicalSystem::dynamics (src/System/src/FloatingBaseSystemDynamics.cpp:
// 226-248 of the reference):
//     m_knownCoefficent.tail(m_actuatedDoFs) += jointTorques;
//     m_generalizedRobotAcceleration = (massMatrix [+ regularization]).llt().solve(m_knownCoefficent);
// for n independent systems of nc = 6 + actuated DoFs unknowns each.  Eigen's LLT (third party, not
// in the reference tree) is the lower Cholesky factorisation A = L L^T read from the LOWER triangle of
// A, followed by the two triangular solves; restated here, not ported.
//
// Warp-level kernel (nc <= 31).  A system is owned by a group of G = 8, 16 or 32 lanes (4, 2, 1
// systems per warp); lane r of the group owns ROW r of A / L in registers and one extra lane owns
// the right-hand side as row nc of the augmented matrix [A b; b^T .], so that the forward
// substitution L y = b comes out of the factorisation itself (y = the last row of the augmented
// factor).  Right-looking, column by column:
//     d = A[j][j] (one shuffle);  rs = 1/sqrt(d);  L[r][j] = A[r][j] * rs;
//     the column goes to shared memory (as row j of L^T, in the tile's unused upper triangle) and
//     every lane updates the rest of its row:  A[r][k] -= L[r][j] * L[k][j],  k > j,
//     L[k][j] read back as a broadcast (one 16-byte shared load per two multiply-adds).
// Lanes never test k <= r: the entries right of a row's diagonal are never read by anybody, so
// they are allowed to hold garbage and the inner loop is a pure LDS + DFMA stream.
// The back substitution L^T x = y walks k = nc-1 .. 0 with x_k broadcast by shuffle and lane r
// reading ITS row of L^T (what the column writes left in the tile).
// Tile rows have an odd pitch (G + 1 doubles), so "lane r reads element k of row r" is free of
// bank conflicts; the global reads are row-wise (lanes = columns <= row): coalesced, and only the
// sectors of the lower triangle are touched.
// Arithmetic deviates from a textbook LLT in one rounding: L[r][j] = s * rsqrt(d) instead of
// s / sqrt(d) (<= 2 ulp per entry; parity tests bound the effect on the solution).  A matrix that is
// not positive definite gives NaN (the reference's Eigen stops the factorisation and solves with the
// partial factor: garbage either way; include/blf_ccm.h states it).
//
// Block-level kernel (32 <= nc <= 128 or forced): one CTA per system, thread r owns row r in
// shared memory, same operations in the same order -- bit-identical to the warp-level kernel where
// both apply (tested) -- three barriers per column; correctness path, not tuned.
#include "dyn_kernels.h"

#include <cstdint>

#include "ccm_ptx.cuh"

namespace blfccm {
namespace {

constexpr int kLltThreads = 128;
constexpr unsigned kFullMask = 0xffffffffu;

template <int G> struct LltTile {
    static constexpr int P = G + 1;          // pitch in doubles, odd
    static constexpr int DOUBLES = G * P;    // G rows (rows >= nc + 1 are only ever touched as garbage)
};

template <int G, int NCMAX, bool REG>
__global__ void __launch_bounds__(kLltThreads)
ccm_llt_solve_kernel(const __grid_constant__ LltArgs a)
{
    static_assert(NCMAX + 1 <= G && (G == 8 || G == 16 || G == 32), "size class");
    constexpr int SPW = 32 / G;
    constexpr int P = LltTile<G>::P;
    extern __shared__ __align__(16) double llt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / G, r = lane % G;
    const int nc = a.nc;
    const long long s0 = (static_cast<long long>(blockIdx.x) * (kLltThreads / 32) + warp) * SPW;
    if (s0 >= a.n) return;   // whole warp
    const long long s = s0 + g;
    const bool valid = s < a.n;
    double* tile = llt_smem + (warp * SPW + g) * LltTile<G>::DOUBLES;

    ptx::grid_dep_wait();

    // ---- lower triangle + right-hand side into the tile (coalesced rows) ------------------------
    {
        const double* M = a.mass + s * nc * nc + r;
#pragma unroll
        for (int i = 0; i < NCMAX; ++i) {
            if (i < nc) {
                if (r <= i) {
                    double v = 1.0;
                    if (valid) {
                        v = __ldcs(M + i * nc);
                        if constexpr (REG) v += __ldg(a.reg + i * nc + r);
                    } else if (r != i) v = 0.0;   // identity for the missing systems of the last warp
                    tile[i * P + r] = v;
                }
            }
        }
        if (r < nc) {
            double v = 0.0;
            if (valid) {
                v = __ldcs(a.known + s * nc + r);
                if (a.tau && r >= 6) v += __ldcs(a.tau + s * (nc - 6) + (r - 6));
            }
            tile[nc * P + r] = v;
        }
    }
    __syncwarp();

    // ---- row r into registers (entries right of the diagonal: whatever the tile holds) ----------
    double row[NCMAX];
#pragma unroll
    for (int k = 0; k < NCMAX; ++k) row[k] = tile[r * P + k];
    __syncwarp();   // the tile's rows are overwritten by columns from here on

    // ---- factorisation, forward substitution riding along in lane nc -----------------------------
    double rdiag = 0.0;   // 1 / L[r][r]
#pragma unroll
    for (int j = 0; j < NCMAX; ++j) {
        if (j < nc) {
            const double d = __shfl_sync(kFullMask, row[j], j, G);
            const double rs = rsqrt(d);
            if (r == j) rdiag = rs;
            const double l = row[j] * rs;
            row[j] = l;
            tile[j * P + r] = l;   // column j = row j of L^T (r > j), y_j at r == nc
            __syncwarp();
            const double nl = -l;
            const double* trow = tile + j * P;
            constexpr int k0c = 0;
            (void)k0c;
            const int k0 = j + 1;
            const int odd = (j * P + k0) & 1;   // 16-byte alignment of the pairs (tile base is aligned)
            if (odd && k0 < NCMAX) row[k0] = fma(nl, trow[k0], row[k0]);
#pragma unroll
            for (int k = k0 + odd; k + 1 < NCMAX; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(trow + k);
                row[k] = fma(nl, v.x, row[k]);
                row[k + 1] = fma(nl, v.y, row[k + 1]);
            }
            if (((NCMAX - (k0 + odd)) & 1) && k0 + odd < NCMAX)
                row[NCMAX - 1] = fma(nl, trow[NCMAX - 1], row[NCMAX - 1]);
        }
    }

    // ---- back substitution: lane r reads its row of L^T, x_k arrives by shuffle -----------------
    double acc = tile[r * P + nc];   // y_r
#pragma unroll
    for (int k = 1; k < NCMAX; ++k) row[k] = tile[r * P + k];
    double x = 0.0;
#pragma unroll
    for (int k = NCMAX - 1; k >= 0; --k) {
        if (k < nc) {
            const double xk = __shfl_sync(kFullMask, acc * rdiag, k, G);
            if (r == k) x = xk;
            if (k > 0) acc = fma(-row[k], xk, acc);
        }
    }
    if (valid && r < nc) a.acc[s * nc + r] = x;
}

// One CTA per system, thread r = row r (r == nc: the right-hand side), everything in shared memory.
constexpr int kLltGenThreads = 160;

__global__ void __launch_bounds__(kLltGenThreads)
ccm_llt_solve_general_kernel(const __grid_constant__ LltArgs a)
{
    extern __shared__ __align__(16) double llt_smem[];
    const int nc = a.nc;
    const int P = (nc + 1) | 1;   // odd pitch >= nc + 1
    double* tile = llt_smem;
    const int r = threadIdx.x;
    const long long s = blockIdx.x;
    ptx::grid_dep_wait();
    {
        const double* M = a.mass + s * nc * nc;
        for (int e = r; e < nc * nc; e += kLltGenThreads) {
            const int i = e / nc, c = e - i * nc;
            if (c <= i) {
                double v = __ldcs(M + e);
                if (a.reg) v += __ldg(a.reg + e);
                tile[i * P + c] = v;
            }
        }
        if (r < nc) {
            double v = __ldcs(a.known + s * nc + r);
            if (a.tau && r >= 6) v += __ldcs(a.tau + s * (nc - 6) + (r - 6));
            tile[nc * P + r] = v;
        }
    }
    __syncthreads();
    double rdiag = 0.0;
    const bool mine = r <= nc;
    for (int j = 0; j < nc; ++j) {
        const double d = tile[j * P + j];
        const double rs = rsqrt(d);
        if (r == j) rdiag = rs;
        double l = 0.0;
        if (mine && r >= j) l = tile[r * P + j] * rs;
        __syncthreads();   // everybody has read the diagonal and its own entry of column j
        if (mine && r > j) tile[j * P + r] = l;
        __syncthreads();
        if (mine && r > j) {
            const double nl = -l;
            const int kend = r < nc ? r : nc - 1;   // the right-hand side row has no diagonal
            for (int k = j + 1; k <= kend; ++k) tile[r * P + k] = fma(nl, tile[j * P + k], tile[r * P + k]);
        }
        __syncthreads();
    }
    double acc = (r < nc) ? tile[r * P + nc] : 0.0;
    double x = 0.0;
    double* xs = tile + nc * P;   // the right-hand side row is free now
    for (int k = nc - 1; k >= 0; --k) {
        if (r == k) {
            x = acc * rdiag;
            xs[k] = x;
        }
        __syncthreads();
        if (r < k) acc = fma(-tile[r * P + k], xs[k], acc);
    }
    if (r < nc) a.acc[s * nc + r] = x;
}

template <typename K>
cudaError_t launch(K kernel, long long grid, int threads, size_t smem, cudaStream_t st, bool pdl,
                   const LltArgs& a)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

template <int G, int NCMAX>
cudaError_t launch_fast(const LltArgs& a, cudaStream_t st, bool pdl)
{
    constexpr int SPW = 32 / G;
    const long long per_block = static_cast<long long>(kLltThreads / 32) * SPW;
    const long long grid = (a.n + per_block - 1) / per_block;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const size_t smem = size_t(kLltThreads / 32) * SPW * LltTile<G>::DOUBLES * sizeof(double);
    return a.reg ? launch(ccm_llt_solve_kernel<G, NCMAX, true>, grid, kLltThreads, smem, st, pdl, a)
                 : launch(ccm_llt_solve_kernel<G, NCMAX, false>, grid, kLltThreads, smem, st, pdl, a);
}

}  // namespace

// size classes of the warp-level kernel: the inner loops are unrolled to NCMAX, so a system of nc
// unknowns pays for the next class up -- classes are at most two apart
#define BLF_LLT_CLASSES(X)                                                                     \
    X(8, 3) X(8, 4) X(8, 5) X(8, 6) X(8, 7)                                                       \
    X(16, 8) X(16, 9) X(16, 10) X(16, 12) X(16, 14) X(16, 15)                                     \
    X(32, 16) X(32, 18) X(32, 20) X(32, 22) X(32, 24) X(32, 26) X(32, 28) X(32, 29) X(32, 30) X(32, 31)

cudaError_t llt_solve_launch(const LltArgs& a, cudaStream_t st, bool pdl, int force_general,
                             int* path_out, int* ncmax_out)
{
    if (a.n <= 0) return cudaSuccess;
    if (!force_general && a.nc <= kLltMaxFast) {
#define BLF_LLT_TRY(G, NCMAX)                         \
    if (a.nc <= NCMAX) {                              \
        if (path_out) *path_out = G;                  \
        if (ncmax_out) *ncmax_out = NCMAX;            \
        return launch_fast<G, NCMAX>(a, st, pdl);     \
    }
        BLF_LLT_CLASSES(BLF_LLT_TRY)
#undef BLF_LLT_TRY
    }
    if (a.nc > kLltMaxCols) return cudaErrorInvalidValue;
    if (a.n > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (path_out) *path_out = 0;
    if (ncmax_out) *ncmax_out = a.nc;
    const int P = (a.nc + 1) | 1;
    const size_t smem = size_t(a.nc + 1) * P * sizeof(double);
    return launch(ccm_llt_solve_general_kernel, a.n, kLltGenThreads, smem, st, pdl, a);
}

}  // namespace blfccm
