// dyn_kernels.cu -- batched mass-matrix solve: the last step of
// System::FloatingBaseDynamicalSystem::dynamics (src/System/src/FloatingBaseSystemDynamics.cpp:
// 226-248 of the reference):
//     m_knownCoefficent.tail(m_actuatedDoFs) += jointTorques;
//     m_generalizedRobotAcceleration = (massMatrix [+ regularization]).llt().solve(m_knownCoefficent);
// for n independent systems of nc = 6 + actuated DoFs unknowns each.  Eigen's LLT (third party, not
// in the reference tree) is the lower Cholesky factorisation A = L L^T read from the LOWER triangle of
// A, followed by the two triangular solves; restated here, not ported.
//
// Warp-level kernel (nc <= 64).  A system is owned by a group of H = 4, 8, 16 or 32 lanes (8, 4, 2 or 1
// systems per warp); lane r of the group owns the ROWS r, r + H, r + 2H, ... of A / L in registers ("slots"),
// and the right-hand side rides along as the last row of the augmented matrix [A b; b^T .] (or,
// where that row would open a slot of its own, as one more column), so that the forward substitution
// L y = b comes out of the factorisation itself (y = the last row of the augmented factor).  Right-looking, column by column:
//     d = A[j][j] (one shuffle);  rs = 1/sqrt(d);  L[i][j] = A[i][j] * rs  for the lane's rows i > j;
//     the column goes through a small shared-memory buffer and every lane updates the rest of its
//     rows:  A[i][k] -= L[i][j] * L[k][j],  k > j,  L[k][j] read back as a broadcast.
// What bounds this kernel is the shared-memory data pipe (one wavefront per broadcast double per
// warp; ncu of the first one-row-per-lane form: 86 % busy at 20 % of HBM) and, with few warps, the
// latency of the column-to-column dependency -- not FP64 issue and not HBM.  A broadcast value read
// once is used for every row the lane owns and for every system of the warp, so several rows per
// lane and several systems per warp divide that traffic.
// Lanes never test k <= i: the entries right of a row's diagonal are never read by anybody, so they
// may hold garbage and the inner loop is a pure LDS + DFMA stream.  The size class N (rows incl. the
// right-hand side) is a template parameter; a smaller system is padded with identity rows (exact
// no-ops: the real entries see the same operations in the same order).
// The back substitution L^T x = y needs COLUMN k of L for x_k -- which is exactly what the lanes of
// a group hold between them (each its rows' entries): per k every lane multiplies its rows' entries
// by its rows' x (the right-hand side row counts as x = -1), shuffles sum over the group (a butterfly
// per unknown on 4-lane groups; on wider ones one reduce-scatter per slot of H unknowns, and inside a
// slot one broadcast per unknown off the slot's diagonal block, kept in shared memory), and the lane
// that owns row k keeps -sum / L[k][k].  So the rows of L never go back to shared memory,
// the tile of a system is only its input triangle, and that tile is refilled for the warp's NEXT
// group of systems (per-thread async copies, LDGSTS: no registers, no waiting) as soon as the rows
// are in registers -- the loads of the next group overlap the whole factorisation of this one.  (A
// first form that loaded, then computed, spent 77 % of its time waiting for its loads; a second
// that kept L^T in shared memory for the back substitution had room for only 6 warps per SM.)
// Shared memory per system: the packed lower triangle of A, N(N+1)/2 doubles (row i at i(i+1)/2;
// rows H apart x 8-byte accesses are free of bank conflicts), the joint torques and two column
// buffers.  The global reads are row-wise (lanes = columns <= row): coalesced, and only the sectors
// of the lower triangle are touched.
// Arithmetic deviates from a textbook LLT in two places: L[i][j] = s * rsqrt(d) instead of
// s / sqrt(d) (<= 2 ulp per entry), and the back substitution's sums are taken in tree order;
// parity tests bound the effect on the solution.  A matrix that is not positive definite gives NaN
// (the reference's Eigen stops the factorisation and solves with the partial factor: garbage either
// way; include/blf_ccm.h states it).
//
// Block-level kernel (65 <= nc <= 128 or forced): one CTA per system, the matrix in shared memory, the
// same factorisation in the same column order with the trailing update spread over the CTA (a warp
// per row), two barriers per column; the path for sizes the register-resident form cannot hold.
#include "dyn_kernels.h"

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "ccm_ptx.cuh"

namespace blfccm {
namespace {

// compile-time loop: f(std::integral_constant<int, 0>) ... f(std::integral_constant<int, N-1>)
template <int N, int I = 0, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<N, I + 1>(f);
    }
}

constexpr int kLltThreads = 64;   // two warps share the regularisation triangle; CTAs per SM from the occupancy API
constexpr unsigned kFullMask = 0xffffffffu;

// shared memory of one system
template <int H, int N> struct LltTile {
    static constexpr int TRI = N * (N + 1) / 2;                        // packed lower triangle + right-hand side row
    static constexpr int TAU = (N - 1 > 6) ? N - 1 - 6 : 0;           // joint torques
    static constexpr int TAU_AT = (TRI + 1) / 2 * 2;
    static constexpr int NP = (N + 1) / 2 * 2;                         // one column buffer
    static constexpr int COL_AT = (TAU_AT + TAU + 1) / 2 * 2;
    // groups of 8, 16 and 32 lanes: the strictly lower part of every H x H diagonal block of L, column by
    // column (H (H - 1) / 2 doubles a block), for the blocked back substitution
    // (measured per class, profiles/r02_llt_variants.log: +6 .. +15 % everywhere except the 24-row class
    // of 8-lane groups, -9 %, which keeps the unblocked form)
#ifdef BLF_LLT_UNBLOCKED
    static constexpr bool BLOCKED = false;
#else
    static constexpr bool BLOCKED = (H == 8 && N != 24) || H == 16 || H == 32;
#endif
    static constexpr int DB = H * (H - 1) / 2;
    static constexpr int DIAG = BLOCKED ? DB * ((N + H - 1) / H) : 0;
    static constexpr int DIAG_AT = COL_AT + 2 * NP;
    static constexpr int STRIDE = (DIAG_AT + DIAG + 15) / 16 * 16 + (H == 8 ? 8 : 4);   // + bank spreading of the groups
    static constexpr int PER_WARP = (32 / H) * STRIDE;
    static constexpr size_t bytes(bool reg)
    {
        return sizeof(double) * (size_t(kLltThreads / 32) * PER_WARP + (reg ? TAU_AT : 0));
    }
    // Bulk-copy route (4-lane classes, exact fit, even order: see ccm_llt_solve_kernel): the tile holds the
    // system's DENSE block, row i at i * (N - 1), then the right-hand side (= row N - 1 of the augmented
    // matrix), the joint torques and the two column buffers; every piece starts on a 16-byte boundary.
    static constexpr int B_RHS_AT = (N - 1) * (N - 1);
    static constexpr int B_TAU_AT = B_RHS_AT + (N - 1);
    static constexpr int B_COL_AT = (B_TAU_AT + TAU + 1) / 2 * 2;
    static constexpr int B_STRIDE = (B_COL_AT + 2 * NP + 15) / 16 * 16 + 4;
    static constexpr int B_PER_WARP = (32 / H) * B_STRIDE;
    static constexpr size_t bytes_bulk(bool reg)
    {
        return sizeof(double) * (size_t(kLltThreads / 32) * B_PER_WARP + (reg ? TAU_AT : 0));
    }
};

// Bulk-copy fill of one system's tile (issued by ONE lane of its group): the dense block, the right-hand
// side and the joint torques are three contiguous runs of global memory, 16-byte aligned when the order is
// even and the arrays are -- three bulk copies (TMA engine) signalled on the warp's mbarrier instead of ~25
// per-thread async copies per lane.  A system past the end of the batch gets the identity with plain stores.
template <int H, int N>
__device__ __forceinline__ void llt_bulk_fill(const LltArgs& a, long long s, double* tile, uint32_t bar)
{
    using T = LltTile<H, N>;
    constexpr int NM = N - 1;
    if (s < a.n) {
        const bool tq = T::TAU > 0 && a.tau != nullptr;
        ptx::mbar_arrive_expect_tx(bar, static_cast<uint32_t>((NM * NM + NM + (tq ? T::TAU : 0)) * 8));
        ptx::bulk_g2s(ptx::smem_addr(tile), a.mass + s * (NM * NM), NM * NM * 8, bar);
        ptx::bulk_g2s(ptx::smem_addr(tile + T::B_RHS_AT), a.known + s * NM, NM * 8, bar);
        if constexpr (T::TAU > 0) {
            if (tq) ptx::bulk_g2s(ptx::smem_addr(tile + T::B_TAU_AT), a.tau + s * T::TAU, T::TAU * 8, bar);
        }
    } else {
        for (int i = 0; i < NM; ++i)
            for (int k = 0; k < NM; ++k) tile[i * NM + k] = (i == k) ? 1.0 : 0.0;
        for (int k = 0; k < NM + T::TAU; ++k) tile[T::B_RHS_AT + k] = 0.0;
        ptx::mbar_arrive(bar);
    }
}

// One piece of a system's input into its tile, per-thread async copies (lanes = columns: coalesced
// rows, only the sectors of the lower triangle are touched).  PIECE i < N-1: row i of the lower
// triangle; PIECE N-1: the right-hand side row and the joint torques.  Padding rows and the missing
// systems of the last group are written as identity / zeros with plain stores.
struct LltSource {
    const double* M;      // mass + s * nc * nc + r
    const double* known;  // known + s * nc + r
    const double* tau;    // tau + s * (nc - 6) + r, or nullptr
    double* t;            // the tile, + r
    uint32_t ts;          // its shared-window address
    int nc, r;
    bool valid;
};

template <int H, int N>
__device__ __forceinline__ LltSource llt_source(const LltArgs& a, long long s, double* tile, int r)
{
    LltSource q;
    q.nc = a.nc;
    q.r = r;
    q.valid = s < a.n;
    q.M = a.mass + s * q.nc * q.nc + r;
    q.known = a.known + s * q.nc + r;
    q.tau = a.tau ? a.tau + s * (q.nc - 6) + r : nullptr;
    q.t = tile + r;
    q.ts = ptx::smem_addr(tile) + r * 8;
    return q;
}

template <int H, int N, int PIECE>
__device__ __forceinline__ void llt_issue_piece(const LltSource& q)
{
    using T = LltTile<H, N>;
    constexpr int NM = N - 1;
    if constexpr (PIECE < NM) {
        constexpr int i = PIECE;
#pragma unroll
        for (int qq = 0; qq * H <= i; ++qq) {
            if (q.r + qq * H <= i) {
                if (q.valid && i < q.nc) ptx::cp_async8(q.ts + (i * (i + 1) / 2 + qq * H) * 8, q.M + i * q.nc + qq * H);
                else q.t[i * (i + 1) / 2 + qq * H] = (q.r + qq * H == i) ? 1.0 : 0.0;
            }
        }
    } else {
#pragma unroll
        for (int qq = 0; qq * H < NM; ++qq) {
            const int c = q.r + qq * H;
            if (c < NM) {
                if (q.valid && c < q.nc) ptx::cp_async8(q.ts + (NM * (NM + 1) / 2 + qq * H) * 8, q.known + qq * H);
                else q.t[NM * (NM + 1) / 2 + qq * H] = 0.0;
            }
        }
        if constexpr (T::TAU > 0) {
            if (q.tau) {
#pragma unroll
                for (int qq = 0; qq * H < NM - 6; ++qq) {
                    const int c = q.r + qq * H;
                    if (c < NM - 6) {
                        if (q.valid && c < q.nc - 6) ptx::cp_async8(q.ts + (T::TAU_AT + qq * H) * 8, q.tau + qq * H);
                        else q.t[T::TAU_AT + qq * H] = 0.0;
                    }
                }
            }
        }
    }
}

template <int H, int N, int PIECE = 0>
__device__ __forceinline__ void llt_issue_all(const LltSource& q)
{
    llt_issue_piece<H, N, PIECE>(q);
    if constexpr (PIECE + 1 < N) llt_issue_all<H, N, PIECE + 1>(q);
}

// Persistent warps: a warp takes groups of 32 / H systems with a grid stride.
// The whole input of the next group through a real call: made with the lane's rows live in registers,
// where the ~80 inlined address computations would push the kernel over 255 registers.
template <int H, int N>
__device__ __noinline__ void llt_issue_all_call(const LltArgs& a, long long s, double* tile, int r)
{
    llt_issue_all<H, N>(llt_source<H, N>(a, s, tile, r));
    ptx::cp_async_commit();
}

// No register cap: asked for 5 CTAs per SM (or __maxnreg__ 184 / 200) the four-rows-per-lane classes
// spill ~0.5-1 KB per thread and lose 16-60 % (profiles/r02_llt_variants.log).
// BULK (4-lane classes only, chosen by the launcher when the order fills the class exactly (nc == N - 1), is
// even, and mass / known / tau are 16-byte aligned): the tiles are dense and refilled by bulk copies
// (llt_bulk_fill) -- with the per-thread copies the refills cost 35-45 % of the kernel's time, a warp's 32 lanes
// touching 8 systems = 8 cache lines per LDGSTS (7 wavefronts of the shared-memory pipe each).
template <int H, int N, bool REG, bool BULK = false>
__global__ void __launch_bounds__(kLltThreads)
ccm_llt_solve_kernel(const __grid_constant__ LltArgs a)
{
    static_assert((H == 4 || H == 8 || H == 16 || H == 32) && N >= 2 && N <= (H == 32 ? 65 : (H == 16 ? 48 : 4 * H)), "size class");
    static_assert(!BULK || (H == 4 && (N - 1) % 2 == 0), "bulk route: 4-lane classes of even order");
    using T = LltTile<H, N>;
    constexpr int NM = N - 1;                 // order of the (identity-padded) matrix
    constexpr int kPerWarp = BULK ? T::B_PER_WARP : T::PER_WARP;
    constexpr int kStride = BULK ? T::B_STRIDE : T::STRIDE;
    constexpr int kRhsAt = BULK ? T::B_RHS_AT : NM * (NM + 1) / 2;
    constexpr int kTauAt = BULK ? T::B_TAU_AT : T::TAU_AT;
    constexpr int kColAt = BULK ? T::B_COL_AT : T::COL_AT;
    auto rowoff = [](int ic) { return BULK ? ic * NM : ic * (ic + 1) / 2; };
    // Where the right-hand side rides: as row NM of the augmented matrix -- free when the last slot
    // has room for it -- or, when the order is a multiple of H and a row more would open a slot of
    // its own (12 unknowns on 4 lanes, 24 on 8: a third more multiply-adds and registers), as one
    // more COLUMN: every row carries its right-hand side entry b_i, updated like any other entry.
    constexpr bool RHSCOL = NM % H == 0;
    constexpr int ROWS = RHSCOL ? NM : N;     // rows held in slots
    constexpr int R = (ROWS + H - 1) / H;     // rows (slots) per lane
    constexpr int SPW = 32 / H;               // systems per warp
    constexpr bool SHORT = H * R > ROWS;      // the last slot has lanes without a row
    constexpr int WPB = kLltThreads / 32;
    extern __shared__ __align__(16) double llt_smem[];
    // the warp index through a broadcast: the compiler then knows it -- and the loop over the warp's
    // groups below -- to be warp-uniform; without it every shuffle in the loop is wrapped in a
    // WARPSYNC.COLLECTIVE / ENDCOLLECTIVE pair (a quarter of the instructions of the 29-unknown kernel)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(kFullMask, threadIdx.x >> 5, 0);
    const int g = lane / H, r = lane % H;
    const int nc = a.nc;
    const long long ngroups = (a.n + SPW - 1) / SPW;
    const long long gstride = static_cast<long long>(gridDim.x) * WPB;
    long long grp = static_cast<long long>(blockIdx.x) * WPB + warp;
    double* const tile = llt_smem + warp * kPerWarp + g * kStride;
    double* const col = tile + kColAt;
    const double* regt = llt_smem + WPB * kPerWarp;
    __shared__ __align__(8) unsigned long long llt_bars[WPB];   // BULK: one mbarrier per warp, one arrival per system
    const uint32_t bar = ptx::smem_addr(&llt_bars[warp]);
    uint32_t phase = 0;
    if constexpr (BULK) {
        if (lane == 0) {
            ptx::mbar_init(bar, SPW);
            ptx::fence_mbar_init();
        }
        __syncwarp();
    }

    ptx::grid_dep_wait();

    if constexpr (REG) {   // lower triangle of the regularisation term, packed like the tiles, once per CTA
        double* w = llt_smem + WPB * kPerWarp;
        for (int i = 0; i < N; ++i)   // rows >= nc (identity padding, the right-hand side row): + 0.0
            for (int c = threadIdx.x; c <= i; c += kLltThreads)
                w[i * (i + 1) / 2 + c] = (i < nc) ? __ldg(a.reg + i * nc + c) : 0.0;
        __syncthreads();
    }
    if (grp >= ngroups) return;

    if constexpr (BULK) {
        if (r == 0) llt_bulk_fill<H, N>(a, grp * SPW + g, tile, bar);
    } else {
        llt_issue_all<H, N>(llt_source<H, N>(a, grp * SPW + g, tile, r));
        ptx::cp_async_commit();
    }
    for (; grp < ngroups; grp += gstride) {
        const long long s = grp * SPW + g;
        const bool valid = s < a.n;
        if constexpr (BULK) {
            ptx::mbar_wait(bar, phase);
            phase ^= 1u;
        } else {
            ptx::cp_async_wait<0>();
            __syncwarp();
        }

        // ---- the lane's rows into registers (entries right of a diagonal: whatever the tile holds:
        //      entries of the same system's later rows) ------------------------------------------------
        double row[R][NM];
        bool rowok[R];
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = r + m * H;
            rowok[m] = !SHORT || m < R - 1 || i < ROWS;
            const int ic = (SHORT && m == R - 1) ? min(i, ROWS - 1) : i;
            const double* src = tile + rowoff(ic);
#pragma unroll
            for (int k = 0; k < NM; ++k)
                if (k < H * (m + 1)) {
                    row[m][k] = src[k];
                    if constexpr (REG) row[m][k] += regt[ic * (ic + 1) / 2 + k];   // M + reg, one rounding (:236-239)
                }
        }
        double b[R], ysave[R];   // RHSCOL: the rows' right-hand side entries, y_i once row i is the pivot
        if constexpr (RHSCOL) {
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const int i = r + m * H;      // SHORT cannot happen here: ROWS is a multiple of H
                b[m] = tile[kRhsAt + i];
                ysave[m] = 0.0;
                if constexpr (T::TAU > 0) {   // known.tail += jointTorques (:226-227)
                    if (a.tau && i >= 6) b[m] += tile[kTauAt + i - 6];
                }
            }
        } else if constexpr (T::TAU > 0) {
            // the right-hand side row: known.tail += jointTorques (:226-227)
            if (a.tau && r == NM % H) {
#pragma unroll
                for (int k = 6; k < NM; ++k) row[R - 1][k] += tile[kTauAt + k - 6];
            }
        }
        __syncwarp();
        // The tile is free: the next group's triangle lands in it while this one is factorised.
        // How the copies are issued is a measured choice per class (tools/micro/llt_bench.cu,
        // profiles/r02_llt_variants.log): groups of 4 lanes -- one piece per column step, inlined;
        // groups of 8 lanes -- all at once through a real call (spread over the steps they cost
        // 13-23 % there, inlined at once they spill).
        constexpr bool SPREAD = H == 4;
#ifdef BLF_LLT_BENCH_NOFILL
        // measurement aid of tools/micro/llt_bench.cu only: every pass factorises the first group's tile again --
        // the kernel without its refills (no LDGSTS, no HBM traffic but the results): the most a cheaper fill could give
        const bool more = false;
#else
        const bool more = grp + gstride < ngroups;
#endif
        const LltSource next = llt_source<H, N>(a, (grp + gstride) * SPW + g, tile, r);
        if constexpr (BULK) {
            if (more && r == 0) llt_bulk_fill<H, N>(a, (grp + gstride) * SPW + g, tile, bar);
        } else if constexpr (SPREAD) {
            if (more) llt_issue_piece<H, N, NM>(next);
        } else {
            if (more) llt_issue_all_call<H, N>(a, (grp + gstride) * SPW + g, tile, r);
        }

        // ---- factorisation; the forward substitution rides along in the last row -----------------
        double rdiag[R];
#pragma unroll
        for (int m = 0; m < R; ++m) rdiag[m] = 0.0;
        static_for<NM>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int mj = j / H, rj = j % H;
            if constexpr (SPREAD && !BULK) {
                if (more) llt_issue_piece<H, N, j>(next);
            }
            const double rs = rsqrt(__shfl_sync(kFullMask, row[mj][j], rj, H));
            if (r == rj) rdiag[mj] = rs;
            double* cb = col + (j & 1) * T::NP;   // cb[i] = L[i][j]; two buffers: one __syncwarp per column
            double nl[R];
#pragma unroll
            for (int m = 0; m < R; ++m) {
                if (m >= mj) {
                    const double l = row[m][j] * rs;
                    row[m][j] = l;   // L stays in the registers of its row's lane (back substitution)
                    nl[m] = -l;
                    const int i = r + m * H;
                    if (i > j && rowok[m]) cb[i] = l;
                    if constexpr (T::BLOCKED) {   // column j of its H x H diagonal block, rows below the diagonal
                        if (m == mj && r > rj && rowok[m])
                            tile[T::DIAG_AT + T::DB * mj + ((H - 1) * rj - rj * (rj - 1) / 2) - rj - 1 + r] = l;
                    }
                }
            }
            if constexpr (RHSCOL) {   // y_j = b_j / L_jj, broadcast like the column
                const double yj = b[mj] * rs;
                if (r == rj) {
                    ysave[mj] = yj;
                    cb[NM] = yj;
                }
            }
            __syncwarp();
            const int k0 = j + 1;
            const int odd = k0 & 1;   // 16-byte alignment of the pairs
            auto upd = [&](int k, double t) {
#pragma unroll
                for (int m = 0; m < R; ++m)
                    if (m >= k / H) row[m][k] = fma(nl[m], t, row[m][k]);
            };
            if (odd && k0 < NM) upd(k0, cb[k0]);
#pragma unroll
            for (int k = k0 + odd; k + 1 < NM; k += 2) {
                const double2 v = *reinterpret_cast<const double2*>(cb + k);
                upd(k, v.x);
                upd(k + 1, v.y);
            }
            if (((NM - (k0 + odd)) & 1) && k0 + odd < NM) upd(NM - 1, cb[NM - 1]);
            if constexpr (RHSCOL) {   // b_i -= L_ij y_j for the rows below the pivot
                const double t = cb[NM];
#pragma unroll
                for (int m = 0; m < R; ++m)
                    if (m >= mj) b[m] = fma(nl[m], t, b[m]);
            }
        });
        if constexpr (SPREAD && !BULK) {
            if (more) ptx::cp_async_commit();
        }

        // ---- back substitution: x_k = (y_k - sum over rows i > k of L[i][k] x_i) / L[k][k]; with the
        //      right-hand side as a row, that row (i = NM) enters the sum with x = -1 and y_k is inside it
        double x[R];
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = 0.0;
        if constexpr (!RHSCOL) {
            if (r == NM % H) x[R - 1] = -1.0;
        }
        if constexpr (T::BLOCKED) {
            // Blocked, H unknowns (one slot) at a time, last slot first.  (a) The part of the sums that
            // comes from the slots already solved is one multiply-add per slot and unknown, summed over
            // the group by a reduce-scatter (H - 1 shuffles for H sums: lane r ends up with the sum of
            // ITS unknown); (b) inside the block the unknowns follow one another with a single
            // broadcast each, lane r reading column r of the block from shared memory (left there
            // during the factorisation).  The unblocked form below spends log2(H) dependent shuffle
            // rounds per unknown: at 29 unknowns 29 x ~150 cycles of latency against 4 x ~500 here.
            __syncwarp();
            // ld[DB mb + kk] = L[H mb + kk][H mb + r]
            const double* ld = tile + T::DIAG_AT + ((H - 2) * r - r * (r - 1) / 2) - 1;
            static_for<R>([&](auto mc) {
                constexpr int mb = R - 1 - decltype(mc)::value;
                double acc = 0.0;
                if constexpr (RHSCOL) acc = ysave[mb];
                if constexpr (mb < R - 1) {
                    auto part = [&](int kk) {   // this lane's share of the sum of unknown H mb + kk
                        double q = 0.0;
#pragma unroll
                        for (int m = mb + 1; m < R; ++m) q = fma(row[m][H * mb + kk], x[m], q);
                        return q;
                    };
                    double p[H / 2];   // the first exchange takes the shares as they are computed
#pragma unroll
                    for (int t = 0; t < H / 2; ++t) {
                        const double lo = part(t), hi = part(t + H / 2);
                        const double keep = (r & (H / 2)) ? hi : lo, send = (r & (H / 2)) ? lo : hi;
                        p[t] = keep + __shfl_xor_sync(kFullMask, send, H / 2, H);
                    }
#pragma unroll
                    for (int bit = H / 4; bit >= 1; bit >>= 1) {
#pragma unroll
                        for (int t = 0; t < bit; ++t) {
                            const double keep = (r & bit) ? p[t + bit] : p[t], send = (r & bit) ? p[t] : p[t + bit];
                            p[t] = keep + __shfl_xor_sync(kFullMask, send, bit, H);
                        }
                    }
                    acc -= p[0];
                }
#pragma unroll
                for (int kk = H - 1; kk >= 0; --kk) {
                    const int i = H * mb + kk;
                    if (i < ROWS) {
                        if (!RHSCOL && i == NM) {   // the right-hand side row: x = -1, nothing to solve
                            if (kk > 0) acc += ld[T::DB * mb + kk];
                        } else {
                            const double xk = __shfl_sync(kFullMask, acc * rdiag[mb], kk, H);
                            if (r == kk) x[mb] = xk;
                            if (kk > 0) acc = fma(-ld[T::DB * mb + kk], xk, acc);
                        }
                    }
                }
            });
        } else {
#pragma unroll
            for (int k = NM - 1; k >= 0; --k) {
                const int mk = k / H, rk = k % H;
                double p = (r > rk) ? row[mk][k] * x[mk] : 0.0;   // rows of slot mk above k (the others: x not yet known)
#pragma unroll
                for (int m = 0; m < R; ++m)
                    if (m > mk) p = fma(row[m][k], x[m], p);
#pragma unroll
                for (int w = 1; w < H; w <<= 1) p += __shfl_xor_sync(kFullMask, p, w, H);
                if constexpr (RHSCOL) {
                    if (r == rk) x[mk] = (ysave[mk] - p) * rdiag[mk];
                } else {
                    if (r == rk) x[mk] = -p * rdiag[mk];
                }
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int i = r + m * H;
            if (valid && i < nc) a.acc[s * nc + i] = x[m];
        }
    }
}

// One CTA per system, everything in shared memory: the augmented matrix (rows 0..nc, row nc = the
// right-hand side) with an odd pitch, the scaled pivot column, 1/L_jj and y.  Per column: every thread
// scales entries of the column (L[i][j], kept in place for the back substitution), then the trailing
// update is spread over the CTA -- a warp per row, lanes over the row's entries -- so no thread walks
// a whole row on its own (the first version did: one dependent load-multiply-store chain per thread,
// 7.7 M systems/s at 64 unknowns).  Same operations on the same operands in the same column order as
// the warp-level kernel; the back substitution subtracts x_k L[k][i] in descending k.
constexpr int kLltGenThreads = 256;

__global__ void __launch_bounds__(kLltGenThreads)
ccm_llt_solve_general_kernel(const __grid_constant__ LltArgs a)
{
    extern __shared__ __align__(16) double llt_smem[];
    const int nc = a.nc;
    const int P = (nc + 1) | 1;            // odd pitch >= nc + 1
    double* tile = llt_smem;               // (nc + 1) x P
    double* colb = tile + (nc + 1) * P;    // nc + 1: L[i][j] of the current column (i == nc: y_j)
    double* rd = colb + (nc + 2);          // nc: 1 / L_jj
    double* ys = rd + (nc + 1);            // nc: y, then x
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int W = kLltGenThreads / 32;
    const long long s = blockIdx.x;
    ptx::grid_dep_wait();
    {
        const double* M = a.mass + s * nc * nc;
        for (int i = warp; i < nc; i += W)          // rows of the lower triangle, lanes = columns
            for (int c = lane; c <= i; c += 32) {
                double v = __ldcs(M + i * nc + c);
                if (a.reg) v += __ldg(a.reg + i * nc + c);
                tile[i * P + c] = v;
            }
        for (int c = t; c < nc; c += kLltGenThreads) {
            double v = __ldcs(a.known + s * nc + c);
            if (a.tau && c >= 6) v += __ldcs(a.tau + s * (nc - 6) + (c - 6));
            tile[nc * P + c] = v;
        }
    }
    __syncthreads();
    for (int j = 0; j < nc; ++j) {
        const double rs = rsqrt(tile[j * P + j]);   // untouched by the updates of earlier columns' barriers
        if (t == 0) rd[j] = rs;
        for (int i = j + 1 + t; i <= nc; i += kLltGenThreads) {
            const double l = tile[i * P + j] * rs;
            tile[i * P + j] = l;
            colb[i] = l;
        }
        __syncthreads();
        for (int i = j + 1 + warp; i <= nc; i += W) {
            const double nl = -colb[i];
            const int kend = i < nc ? i : nc - 1;   // the right-hand side row has no diagonal
            double* ri = tile + i * P;
            for (int k = j + 1 + lane; k <= kend; k += 32) ri[k] = fma(nl, colb[k], ri[k]);
        }
        __syncthreads();
    }
    for (int i = t; i < nc; i += kLltGenThreads) ys[i] = tile[nc * P + i];   // y
    __syncthreads();
    for (int k = nc - 1; k >= 0; --k) {
        if (t == 0) ys[k] = ys[k] * rd[k];          // x_k
        __syncthreads();
        const double xk = ys[k];
        for (int i = t; i < k; i += kLltGenThreads) ys[i] = fma(-tile[k * P + i], xk, ys[i]);
        __syncthreads();
    }
    for (int i = t; i < nc; i += kLltGenThreads) a.acc[s * nc + i] = ys[i];
}

template <typename K>
cudaError_t launch(K kernel, long long grid, int threads, size_t smem, cudaStream_t st, bool pdl,
                   const LltArgs& a, bool set_attr = true)
{
    if (set_attr && smem > 48 * 1024) {
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

template <int H, int N, bool REG, bool BULK = false>
cudaError_t launch_fast_k(const LltArgs& a, cudaStream_t st, bool pdl)
{
    auto kernel = ccm_llt_solve_kernel<H, N, REG, BULK>;
    const size_t smem = BULK ? LltTile<H, N>::bytes_bulk(REG) : LltTile<H, N>::bytes(REG);
    // resident CTAs of this instantiation per device (CTAs per SM x SMs), filled on first use; an
    // atomic because distinct handles may make their first call from different host threads
    static std::atomic<int> resident[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    int cap = resident[dev].load(std::memory_order_acquire);
    if (cap == 0) {
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return e;
        }
        int v = 0, n = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kLltThreads, smem);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (v < 1) return cudaErrorLaunchOutOfResources;
        cap = v * n;
        resident[dev].store(cap, std::memory_order_release);
    }
    constexpr int SPW = 32 / H;
    const long long groups = (a.n + SPW - 1) / SPW;
    const long long want = (groups + kLltThreads / 32 - 1) / (kLltThreads / 32);
    const long long grid = std::min<long long>(want, cap);
    return launch(kernel, grid, kLltThreads, smem, st, pdl, a, false);
}

template <int H, int N>
cudaError_t launch_fast(const LltArgs& a, cudaStream_t st, bool pdl)
{
    // the bulk-copy route: 6, 8, 10, 12 and 14 unknowns (the 4-lane classes an even order fills exactly), arrays on
    // 16-byte boundaries; BLF_CCM_TUNE_LLT_NO_BULK=1 keeps the per-thread copies (A/B, tests of both routes)
    if constexpr (H == 4 && (N == 7 || N == 9 || N == 11 || N == 13 || N == 15)) {
        static const bool no_bulk = [] {
            const char* e = std::getenv("BLF_CCM_TUNE_LLT_NO_BULK");
            return e && e[0] == '1';
        }();
        auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
        if (!no_bulk && a.nc == N - 1 && al16(a.mass) && al16(a.known) && al16(a.tau))
            return a.reg ? launch_fast_k<H, N, true, true>(a, st, pdl) : launch_fast_k<H, N, false, true>(a, st, pdl);
    }
    return a.reg ? launch_fast_k<H, N, true>(a, st, pdl) : launch_fast_k<H, N, false>(a, st, pdl);
}

}  // namespace

// Size classes of the warp-level kernel, X(lanes per system, rows incl. the right-hand side): a
// system of nc unknowns runs in the first class with N >= nc + 1, padded with identity rows.  The
// classes are compiled in two translation units side by side (this file, and this file again
// included by dyn_kernels_wide.cu with BLF_LLT_TU_WIDE defined: the 16- and 32-lane classes), which
// halves the build's wall time; tools/micro/llt_bench.cu includes this file once with a short list.
#if defined(BLF_LLT_BENCH_CLASSES)
#if defined(BLF_LLT_ONLY_WIDE)
#define BLF_LLT_CLASSES(X) X(32, 64)
#elif defined(BLF_LLT_ONLY_29_H16)
#define BLF_LLT_CLASSES(X) X(16, 30)
#elif defined(BLF_LLT_ONLY_29_H32)
#define BLF_LLT_CLASSES(X) X(32, 30)
#elif defined(BLF_LLT_ONLY_MID)
#define BLF_LLT_CLASSES(X) X(8, 17) X(8, 19) X(8, 20) X(8, 22) X(8, 24) X(8, 26) X(8, 28)
#else
#define BLF_LLT_CLASSES(X) X(4, 7) X(4, 13) X(8, 19) X(8, 24) X(8, 25) X(8, 30) X(16, 39)
#endif
#define BLF_LLT_NARROW_MAX kLltMaxFast
#elif defined(BLF_LLT_TU_WIDE)
#define BLF_LLT_CLASSES(X)                                                                               \
    X(16, 33) X(16, 34) X(16, 36) X(16, 37) X(16, 39) X(16, 40) X(16, 42) X(16, 45) X(16, 48)          \
    X(32, 52) X(32, 56) X(32, 60) X(32, 64) X(32, 65)
#else
#define BLF_LLT_CLASSES(X)                                                      \
    X(4, 4) X(4, 5) X(4, 7) X(4, 8) X(4, 9) X(4, 10) X(4, 11) X(4, 13) X(4, 14) X(4, 15) X(4, 16)   \
    X(8, 17) X(8, 19) X(8, 20) X(8, 22) X(8, 24) X(8, 25) X(8, 26) X(8, 28) X(8, 30) X(8, 31) X(8, 32)
#define BLF_LLT_NARROW_MAX 31
#endif

#define BLF_LLT_TRY(H, N)                          \
    if (a.nc < N) {                                \
        if (path_out) *path_out = H;               \
        if (ncmax_out) *ncmax_out = N - 1;         \
        return launch_fast<H, N>(a, st, pdl);      \
    }

#ifdef BLF_LLT_TU_WIDE
// the 16- and 32-lane classes (32 .. 64 unknowns); called by llt_solve_launch of the other unit
cudaError_t llt_solve_launch_wide(const LltArgs& a, cudaStream_t st, bool pdl, int* path_out, int* ncmax_out)
{
    BLF_LLT_CLASSES(BLF_LLT_TRY)
    return cudaErrorInvalidValue;
}
#else
cudaError_t llt_solve_launch_wide(const LltArgs& a, cudaStream_t st, bool pdl, int* path_out, int* ncmax_out);

cudaError_t llt_solve_launch(const LltArgs& a, cudaStream_t st, bool pdl, int force_general,
                             int* path_out, int* ncmax_out)
{
    if (a.n <= 0) return cudaSuccess;
    if (!force_general && a.nc <= kLltMaxFast) {
        if (a.nc <= BLF_LLT_NARROW_MAX) {
            BLF_LLT_CLASSES(BLF_LLT_TRY)
        }
#ifndef BLF_LLT_BENCH_CLASSES
        return llt_solve_launch_wide(a, st, pdl, path_out, ncmax_out);
#endif
    }
    if (a.nc > kLltMaxCols) return cudaErrorInvalidValue;
    if (a.n > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (path_out) *path_out = 0;
    if (ncmax_out) *ncmax_out = a.nc;
    const int P = (a.nc + 1) | 1;
    const size_t smem = (size_t(a.nc + 1) * P + 3 * size_t(a.nc) + 4) * sizeof(double);
    return launch(ccm_llt_solve_general_kernel, a.n, kLltGenThreads, smem, st, pdl, a);
}
#endif
#undef BLF_LLT_TRY

}  // namespace blfccm
