// standin_selftest.cpp -- C entry points that exercise the STAND-IN Eigen / iDynTree headers directly,
// so that tests/test_standin_headers.py can compare each operation with numpy.  TEST INFRASTRUCTURE
// ONLY (see ../README.md): the stand-ins carry the "parity pinned against the reference's sources"
// claim, so their own semantics (storage orders, maps, block views, products, inverses, aliasing)
// are checked independently of the reference's code.
#include <cstring>

#include <Eigen/Dense>
#include <iDynTree/Core/EigenHelpers.h>

using Eigen::Index;
using RowMajorDyn = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;

extern "C" {

// C = A(m x k) * B(k x n), all row-major buffers, through dynamic row-major maps
void st_matmul(int m, int k, int n, const double* a, const double* b, double* c)
{
    Eigen::Map<const RowMajorDyn> A(a, m, k), B(b, k, n);
    Eigen::Map<RowMajorDyn> C(c, m, n);
    C = A * B;
}

// y = (s * A) * A * x + t * (A^T * x) for fixed 3x3 A given ROW-major, via iDynTree types and toEigen
void st_fixed3_chain(const double a_rowmajor[9], const double x[3], double s, double t, double y[3])
{
    iDynTree::Matrix3x3 A(a_rowmajor, 3, 3);
    iDynTree::Vector3 X(x, 3), Y;
    iDynTree::toEigen(Y) = s * iDynTree::toEigen(A) * iDynTree::toEigen(A) * iDynTree::toEigen(X)
                           + t * (iDynTree::toEigen(A).transpose() * iDynTree::toEigen(X));
    std::memcpy(y, Y.data(), 3 * sizeof(double));
}

// column-major fixed Matrix3d filled by (r, c), inverse() -> row-major out
void st_inverse3(const double a_rowmajor[9], double out_rowmajor[9])
{
    Eigen::Matrix3d A;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) A(r, c) = a_rowmajor[3 * r + c];
    const Eigen::Matrix3d inv = A.inverse();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out_rowmajor[3 * r + c] = inv(r, c);
}

// dynamic inverse (partial-pivot LU) and LLT solve, row-major buffers
void st_inverse_dyn(int n, const double* a, double* out)
{
    Eigen::Map<const RowMajorDyn> A(a, n, n);
    Eigen::Map<RowMajorDyn> O(out, n, n);
    O = A.inverse();
}
void st_llt_solve(int n, const double* a, const double* b, double* x)
{
    Eigen::Map<const RowMajorDyn> A(a, n, n);
    Eigen::Map<const Eigen::VectorXd> B(b, n);
    Eigen::Map<Eigen::VectorXd> X(x, n);
    X = A.llt().solve(B);
}

// skew(v), cross, colwise().cross: out[0..8] = skew(v) row-major, out[9..11] = v x w,
// out[12..20] = (R.colwise().cross(w)) row-major with R given row-major
void st_cross_ops(const double v[3], const double w[3], const double r_rowmajor[9], double out[21])
{
    Eigen::Vector3d V(v[0], v[1], v[2]), W(w[0], w[1], w[2]);
    const auto S = iDynTree::skew(V);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] = S(r, c);
    const Eigen::Vector3d VW = V.cross(W);
    for (int i = 0; i < 3; ++i) out[9 + i] = VW(i);
    Eigen::Matrix3d R;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R(r, c) = r_rowmajor[3 * r + c];
    const Eigen::Matrix3d CW = R.colwise().cross(W);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[12 + 3 * r + c] = CW(r, c);
}

// block views on a row-major 6x6 (iDynTree::Matrix6x6) and a 6-vector, the forms the reference uses:
//   g.topLeftCorner(3,3).diagonal().array() = d;  g.bottomRightCorner(3,3) = B (3x3 row-major in);
//   f.head(3) = h; f.tail(3) = t;  regressor(6x2).topRightCorner<3,1>() = h; bottomLeftCorner<3,1>() = t
void st_block_writes(double d, const double b_rowmajor[9], const double h[3], const double t[3],
                     double g_out[36], double f_out[6], double reg_out[12])
{
    iDynTree::Matrix6x6 g;
    g.zero();
    iDynTree::Vector6 f;
    f.zero();
    iDynTree::MatrixDynSize reg(6, 2);
    iDynTree::Matrix3x3 B(b_rowmajor, 3, 3);
    iDynTree::Vector3 H(h, 3), T(t, 3);
    auto G = iDynTree::toEigen(g);
    G.topLeftCorner(3, 3).diagonal().array() = d;
    G.bottomRightCorner(3, 3) = iDynTree::toEigen(B);
    auto F = iDynTree::toEigen(f);
    F.head(3) = iDynTree::toEigen(H);
    F.tail(3) = iDynTree::toEigen(T);
    auto Rg = iDynTree::toEigen(reg);
    Rg.topRightCorner<3, 1>() = iDynTree::toEigen(H);
    Rg.bottomLeftCorner<3, 1>() = iDynTree::toEigen(T);
    std::memcpy(g_out, g.data(), sizeof(double) * 36);
    std::memcpy(f_out, f.data(), sizeof(double) * 6);
    std::memcpy(reg_out, reg.data(), sizeof(double) * 12);
}

// aliasing: x = x + K * (z - Y * x) and P = (P - K * Y * P) / lambda with maps over the SAME buffers
// (the reference's RecursiveLeastSquare.cpp:125-130), row-major dynamic
void st_aliasing_update(int p, int m, const double* K, const double* Y, const double* z, double lambda,
                        double* x, double* P)
{
    Eigen::Map<const RowMajorDyn> Km(K, p, m), Ym(Y, m, p);
    Eigen::Map<const Eigen::VectorXd> Z(z, m);
    Eigen::Map<Eigen::VectorXd> X(x, p);
    Eigen::Map<RowMajorDyn> Pm(P, p, p);
    X = X + Km * (Z - Ym * X);
    Pm = (Pm - Km * Ym * Pm) / lambda;
}

// diag: out(n x n row-major) = v.asDiagonal()
void st_as_diagonal(int n, const double* v, double* out)
{
    Eigen::Map<const Eigen::VectorXd> V(v, n);
    Eigen::Map<RowMajorDyn> O(out, n, n);
    O = V.asDiagonal();
}

// Rotation::RPY and AngularMotionVector3::exp, row-major out
void st_rpy(double r, double p, double y, double out[9])
{
    const iDynTree::Rotation R = iDynTree::Rotation::RPY(r, p, y);
    std::memcpy(out, R.data(), sizeof(double) * 9);
}
void st_exp(const double w[3], double out[9])
{
    iDynTree::AngularMotionVector3 v(w[0], w[1], w[2]);
    const iDynTree::Rotation R = v.exp();
    std::memcpy(out, R.data(), sizeof(double) * 9);
}

} // extern "C"
