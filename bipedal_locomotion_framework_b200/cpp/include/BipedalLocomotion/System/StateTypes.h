/**
 * @file StateTypes.h
 * Small dense value types for the System facade.  The reference uses Eigen::Vector3d,
 * Eigen::Matrix3d, Eigen::Matrix<double,6,1> and Eigen::VectorXd here
 * (src/System/include/BipedalLocomotion/System/FloatingBaseSystemKinematics.h:31-33); Eigen is a
 * third-party dependency that is not available in this build, so these stand-ins provide the
 * storage (Matrix3d is ROW-major here) and the two operations the integrators need
 * (`x += dx * dT`, ForwardEuler.h:50).
 *
 * With Eigen available define BLF_HAVE_EIGEN (CMake: -DFRAMEWORK_USE_Eigen=ON): the four names
 * then ARE the reference's Eigen types (Matrix3d column-major), so code written against the
 * reference's System classes -- its own IntegratorTest.cpp included -- compiles unchanged.  The
 * facade never assumes a storage order: rotations cross the C ABI through toRowMajor/fromRowMajor.
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_STATE_TYPES_H
#define BIPEDAL_LOCOMOTION_SYSTEM_STATE_TYPES_H

#if defined(BLF_HAVE_EIGEN)

#include <Eigen/Dense>

namespace BipedalLocomotion
{
namespace System
{
using Vector3d = Eigen::Vector3d;
using Vector6d = Eigen::Matrix<double, 6, 1>;
using Matrix3d = Eigen::Matrix3d;
using VectorXd = Eigen::VectorXd;
} // namespace System
} // namespace BipedalLocomotion

#else // bundled value types

#include <array>
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace BipedalLocomotion
{
namespace System
{

template <std::size_t N> struct FixedVector
{
    std::array<double, N> v{};

    FixedVector() = default;
    FixedVector(std::initializer_list<double> l)
    {
        std::size_t i = 0;
        for (double x : l)
            if (i < N) v[i++] = x;
    }
    double& operator()(std::size_t i) { return v[i]; }
    const double& operator()(std::size_t i) const { return v[i]; }
    double& operator[](std::size_t i) { return v[i]; }
    const double& operator[](std::size_t i) const { return v[i]; }
    double* data() { return v.data(); }
    const double* data() const { return v.data(); }
    static constexpr std::size_t size() { return N; }
    void setZero() { v.fill(0.0); }
    FixedVector operator*(double s) const
    {
        FixedVector r;
        for (std::size_t i = 0; i < N; ++i) r.v[i] = v[i] * s;
        return r;
    }
    FixedVector& operator+=(const FixedVector& o)
    {
        for (std::size_t i = 0; i < N; ++i) v[i] = v[i] + o.v[i];
        return *this;
    }
};

using Vector3d = FixedVector<3>;
using Vector6d = FixedVector<6>;

/** 3x3, row-major storage (element (r,c) at 3*r + c). */
struct Matrix3d : FixedVector<9>
{
    Matrix3d() = default;
    double& operator()(std::size_t r, std::size_t c) { return v[3 * r + c]; }
    const double& operator()(std::size_t r, std::size_t c) const { return v[3 * r + c]; }
    void setIdentity()
    {
        setZero();
        v[0] = v[4] = v[8] = 1.0;
    }
    static Matrix3d Identity()
    {
        Matrix3d m;
        m.setIdentity();
        return m;
    }
    Matrix3d operator*(double s) const
    {
        Matrix3d r;
        for (std::size_t i = 0; i < 9; ++i) r.v[i] = v[i] * s;
        return r;
    }
    Matrix3d& operator+=(const Matrix3d& o)
    {
        for (std::size_t i = 0; i < 9; ++i) v[i] = v[i] + o.v[i];
        return *this;
    }
};

/** Dynamic-size vector. */
struct VectorXd
{
    std::vector<double> v;

    VectorXd() = default;
    explicit VectorXd(std::size_t n) : v(n, 0.0) {}
    VectorXd(std::initializer_list<double> l) : v(l) {}
    double& operator()(std::size_t i) { return v[i]; }
    const double& operator()(std::size_t i) const { return v[i]; }
    double& operator[](std::size_t i) { return v[i]; }
    const double& operator[](std::size_t i) const { return v[i]; }
    double* data() { return v.data(); }
    const double* data() const { return v.data(); }
    std::size_t size() const { return v.size(); }
    void resize(std::size_t n) { v.resize(n, 0.0); }
    void setZero() { v.assign(v.size(), 0.0); }
    VectorXd operator*(double s) const
    {
        VectorXd r(v.size());
        for (std::size_t i = 0; i < v.size(); ++i) r.v[i] = v[i] * s;
        return r;
    }
    VectorXd& operator+=(const VectorXd& o)
    {
        for (std::size_t i = 0; i < v.size() && i < o.v.size(); ++i) v[i] = v[i] + o.v[i];
        return *this;
    }
};

} // namespace System
} // namespace BipedalLocomotion

#endif // BLF_HAVE_EIGEN

namespace BipedalLocomotion
{
namespace System
{
/** The C ABI takes rotations as 9 doubles, row-major, whatever Matrix3d's own storage order is. */
inline void toRowMajor(const Matrix3d& m, double out[9])
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out[3 * r + c] = m(r, c);
}
inline void fromRowMajor(const double in[9], Matrix3d& m)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) m(r, c) = in[3 * r + c];
}
} // namespace System
} // namespace BipedalLocomotion

#endif // BIPEDAL_LOCOMOTION_SYSTEM_STATE_TYPES_H
