/**
 * @file StateTypes.h
 * Small dense value types for the System facade.  The reference uses Eigen::Vector3d,
 * Eigen::Matrix3d, Eigen::Matrix<double,6,1> and Eigen::VectorXd here
 * (src/System/include/BipedalLocomotion/System/FloatingBaseSystemKinematics.h:31-33); Eigen is a
 * third-party dependency that is not available in this build, so these stand-ins provide the
 * storage layout (Matrix3d is ROW-major here: it is handed to the C ABI as is) and the two
 * operations the integrators need (`x += dx * dT`, ForwardEuler.h:50).
 */
#ifndef BIPEDAL_LOCOMOTION_SYSTEM_STATE_TYPES_H
#define BIPEDAL_LOCOMOTION_SYSTEM_STATE_TYPES_H

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

#endif // BIPEDAL_LOCOMOTION_SYSTEM_STATE_TYPES_H
