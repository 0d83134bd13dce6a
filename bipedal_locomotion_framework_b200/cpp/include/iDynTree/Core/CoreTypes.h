// CoreTypes.h -- minimal stand-ins for the iDynTree core types the ContactModels interface uses.
//
// iDynTree is a third-party dependency of the reference (>= 0.11.105, CI pin v1.1.0) that is not
// vendored and not installed here.  These classes reproduce only the API SHAPE and the MEMORY
// LAYOUT the contact-model path relies on, so that (a) the facade has the reference's signatures
// and (b) arrays of these objects can be handed to the C ABI unchanged:
//   Twist / Wrench / SpatialAcc / Vector6   6 doubles, linear(3) then angular(3)      48 bytes
//   Transform                              Position(3) then Rotation(3x3 row-major)  96 bytes
//   Matrix6x6                              36 doubles row-major                      288 bytes
// When the real iDynTree is available, define BLF_HAVE_IDYNTREE and include its headers instead;
// ContactModelTypes.h static_asserts the sizes either way.
#ifndef BLF_IDYNTREE_CORE_TYPES_SHIM_H
#define BLF_IDYNTREE_CORE_TYPES_SHIM_H

#include <cmath>
#include <cstddef>
#include <vector>

namespace iDynTree
{

template <unsigned N> class VectorFixSize
{
protected:
    double m_data[N];

public:
    VectorFixSize() { zero(); }
    VectorFixSize(const double* in, std::size_t n)
    {
        for (unsigned i = 0; i < N; ++i) m_data[i] = i < n ? in[i] : 0.0;
    }
    double& operator()(std::size_t i) { return m_data[i]; }
    const double& operator()(std::size_t i) const { return m_data[i]; }
    double& operator[](std::size_t i) { return m_data[i]; }
    const double& operator[](std::size_t i) const { return m_data[i]; }
    double* data() { return m_data; }
    const double* data() const { return m_data; }
    std::size_t size() const { return N; }
    void zero()
    {
        for (unsigned i = 0; i < N; ++i) m_data[i] = 0.0;
    }
};

using Vector2 = VectorFixSize<2>;
using Vector3 = VectorFixSize<3>;
using Vector6 = VectorFixSize<6>;

class Position : public Vector3
{
public:
    Position() = default;
    Position(double x, double y, double z)
    {
        m_data[0] = x;
        m_data[1] = y;
        m_data[2] = z;
    }
};
using Force = Vector3;
using Torque = Vector3;
using LinVelocity = Vector3;
using AngVelocity = Vector3;

template <unsigned R, unsigned C> class MatrixFixSize
{
protected:
    double m_data[R * C]; // row-major

public:
    MatrixFixSize() { zero(); }
    double& operator()(std::size_t r, std::size_t c) { return m_data[r * C + c]; }
    const double& operator()(std::size_t r, std::size_t c) const { return m_data[r * C + c]; }
    double* data() { return m_data; }
    const double* data() const { return m_data; }
    std::size_t rows() const { return R; }
    std::size_t cols() const { return C; }
    void zero()
    {
        for (unsigned i = 0; i < R * C; ++i) m_data[i] = 0.0;
    }
};

using Matrix3x3 = MatrixFixSize<3, 3>;
using Matrix6x6 = MatrixFixSize<6, 6>;

class Rotation : public Matrix3x3
{
public:
    Rotation() { *this = Identity(); }
    static Rotation Identity()
    {
        Rotation r(0);
        r(0, 0) = r(1, 1) = r(2, 2) = 1.0;
        return r;
    }
    /** Rz(yaw) * Ry(pitch) * Rx(roll) */
    static Rotation RPY(double roll, double pitch, double yaw)
    {
        const double cr = std::cos(roll), sr = std::sin(roll);
        const double cp = std::cos(pitch), sp = std::sin(pitch);
        const double cy = std::cos(yaw), sy = std::sin(yaw);
        Rotation r(0);
        r(0, 0) = cy * cp; r(0, 1) = cy * sp * sr - sy * cr; r(0, 2) = cy * sp * cr + sy * sr;
        r(1, 0) = sy * cp; r(1, 1) = sy * sp * sr + cy * cr; r(1, 2) = sy * sp * cr - cy * sr;
        r(2, 0) = -sp;     r(2, 1) = cp * sr;                r(2, 2) = cp * cr;
        return r;
    }
    Rotation operator*(const Rotation& o) const
    {
        Rotation r(0);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                r(i, j) = (*this)(i, 0) * o(0, j) + (*this)(i, 1) * o(1, j) + (*this)(i, 2) * o(2, j);
        return r;
    }

private:
    explicit Rotation(int) { zero(); }
};

/** Rotation vector; exp() is Rodrigues' formula. */
class AngularMotionVector3 : public Vector3
{
public:
    Rotation exp() const
    {
        const double x = m_data[0], y = m_data[1], z = m_data[2];
        const double th = std::sqrt(x * x + y * y + z * z);
        Rotation r = Rotation::Identity();
        if (th < 1e-300) return r;
        const double a = std::sin(th) / th, b = (1.0 - std::cos(th)) / (th * th);
        const double K[3][3] = {{0, -z, y}, {z, 0, -x}, {-y, x, 0}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double kk = 0;
                for (int k = 0; k < 3; ++k) kk += K[i][k] * K[k][j];
                r(i, j) += a * K[i][j] + b * kk;
            }
        return r;
    }
};

class Transform
{
    Position m_pos;
    Rotation m_rot;

public:
    Transform() = default;
    Transform(const Rotation& r, const Position& p) : m_pos(p), m_rot(r) {}
    static Transform Identity() { return Transform(); }
    const Position& getPosition() const { return m_pos; }
    const Rotation& getRotation() const { return m_rot; }
    void setPosition(const Position& p) { m_pos = p; }
    void setRotation(const Rotation& r) { m_rot = r; }
};

/** linear(3) then angular(3) */
class SpatialVector6 : public Vector6
{
public:
    static SpatialVector6 Zero() { return SpatialVector6(); }
    Vector3& getLinearVec3() { return *reinterpret_cast<Vector3*>(m_data); }
    const Vector3& getLinearVec3() const { return *reinterpret_cast<const Vector3*>(m_data); }
    Vector3& getAngularVec3() { return *reinterpret_cast<Vector3*>(m_data + 3); }
    const Vector3& getAngularVec3() const { return *reinterpret_cast<const Vector3*>(m_data + 3); }
};

class Twist : public SpatialVector6
{
public:
    static Twist Zero() { return Twist(); }
};
class Wrench : public SpatialVector6
{
public:
    static Wrench Zero() { return Wrench(); }
};
class SpatialAcc : public SpatialVector6
{
};

class VectorDynSize
{
    std::vector<double> m_data;

public:
    VectorDynSize() = default;
    explicit VectorDynSize(std::size_t n) : m_data(n, 0.0) {}
    void resize(std::size_t n) { m_data.resize(n, 0.0); }
    void zero() { m_data.assign(m_data.size(), 0.0); }
    double& operator()(std::size_t i) { return m_data[i]; }
    const double& operator()(std::size_t i) const { return m_data[i]; }
    double& operator[](std::size_t i) { return m_data[i]; }
    const double& operator[](std::size_t i) const { return m_data[i]; }
    double* data() { return m_data.data(); }
    const double* data() const { return m_data.data(); }
    std::size_t size() const { return m_data.size(); }
};

class MatrixDynSize
{
    std::vector<double> m_data; // row-major
    std::size_t m_rows{0}, m_cols{0};

public:
    MatrixDynSize() = default;
    MatrixDynSize(std::size_t r, std::size_t c) { resize(r, c); }
    void resize(std::size_t r, std::size_t c)
    {
        m_rows = r;
        m_cols = c;
        m_data.assign(r * c, 0.0);
    }
    void zero() { m_data.assign(m_data.size(), 0.0); }
    double& operator()(std::size_t r, std::size_t c) { return m_data[r * m_cols + c]; }
    const double& operator()(std::size_t r, std::size_t c) const { return m_data[r * m_cols + c]; }
    double* data() { return m_data.data(); }
    const double* data() const { return m_data.data(); }
    std::size_t rows() const { return m_rows; }
    std::size_t cols() const { return m_cols; }
};

} // namespace iDynTree

#endif // BLF_IDYNTREE_CORE_TYPES_SHIM_H
