// Stand-in for the SUBSET of iDynTree::Core that the reference's hot-path sources and tests use.
//
// TEST INFRASTRUCTURE ONLY (see oracle/refbuild/README.md).  iDynTree (>= 0.11.105, CI pin v1.1.0:
// cmake/BipedalLocomotionFrameworkFindDependencies.cmake:133, .github/workflows/ci.yml:15) is a
// third-party dependency of the reference that is absent from this image.  This header supplies the
// storage types and helpers from iDynTree's published interface so that the reference's own .cpp
// files compile unmodified:
//   VectorFixSize / VectorDynSize / MatrixFixSize / MatrixDynSize   (row-major storage),
//   Position, Rotation (RPY = Rz(yaw) Ry(pitch) Rx(roll)), Transform = {Position, Rotation},
//   Twist / SpatialAcc / Wrench = {linear 3-vector, angular 3-vector},
//   AngularMotionVector3::exp (Rodrigues), Span / make_span (GSL-style),
//   EigenHelpers: toEigen(...) maps and skew(v) = [[0,-v2,v1],[v2,0,-v0],[-v1,v0,0]].
// Written from the interface, not from iDynTree's sources.
#ifndef BLF_REFBUILD_STANDIN_IDYNTREE_CORE
#define BLF_REFBUILD_STANDIN_IDYNTREE_CORE

#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iterator>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include <Eigen/Core>

namespace iDynTree
{

// ------------------------------------------------------------------------------------------------
// Span (GSL-style view of contiguous memory)
// ------------------------------------------------------------------------------------------------
template <class ElementType> class Span
{
public:
    using element_type = ElementType;
    using value_type = std::remove_cv_t<ElementType>;
    using index_type = std::ptrdiff_t;
    using size_type = index_type;
    using pointer = element_type*;
    using reference = element_type&;
    using iterator = pointer;
    using const_iterator = const element_type*;
    using reverse_iterator = std::reverse_iterator<iterator>;
    using const_reverse_iterator = std::reverse_iterator<const_iterator>;

    constexpr Span() noexcept : m_p(nullptr), m_n(0) {}
    constexpr Span(pointer p, index_type n) : m_p(p), m_n(n) {}
    constexpr Span(pointer first, pointer last) : m_p(first), m_n(last - first) {}
    template <std::size_t N> constexpr Span(element_type (&arr)[N]) noexcept : m_p(arr), m_n(N) {}
    template <std::size_t N, class U = value_type>
    constexpr Span(std::array<U, N>& a) noexcept : m_p(a.data()), m_n(N)
    {
    }
    template <std::size_t N, class U = value_type,
              class = std::enable_if_t<std::is_const<element_type>::value, U>>
    constexpr Span(const std::array<U, N>& a) noexcept : m_p(a.data()), m_n(N)
    {
    }
    // any container with data() and size() whose elements convert without adjustment
    template <class Container,
              class = std::enable_if_t<
                  !std::is_array<Container>::value
                  && std::is_convertible<decltype(std::declval<Container&>().data()), pointer>::value
                  && std::is_convertible<decltype(std::declval<Container&>().size()), index_type>::value>>
    constexpr Span(Container& c) : m_p(c.data()), m_n(static_cast<index_type>(c.size()))
    {
    }
    template <class Container,
              class = std::enable_if_t<
                  std::is_const<element_type>::value && !std::is_array<Container>::value
                  && std::is_convertible<decltype(std::declval<const Container&>().data()), pointer>::value
                  && std::is_convertible<decltype(std::declval<const Container&>().size()), index_type>::value>>
    constexpr Span(const Container& c) : m_p(c.data()), m_n(static_cast<index_type>(c.size()))
    {
    }
    template <class U, class = std::enable_if_t<std::is_convertible<U (*)[], element_type (*)[]>::value>>
    constexpr Span(const Span<U>& o) : m_p(o.data()), m_n(o.size())
    {
    }
    constexpr Span(const Span&) noexcept = default;
    constexpr Span& operator=(const Span&) noexcept = default;

    constexpr index_type size() const noexcept { return m_n; }
    constexpr index_type size_bytes() const noexcept { return m_n * index_type(sizeof(element_type)); }
    constexpr bool empty() const noexcept { return m_n == 0; }
    constexpr pointer data() const noexcept { return m_p; }
    constexpr reference operator[](index_type i) const { return m_p[i]; }
    constexpr reference operator()(index_type i) const { return m_p[i]; }
    constexpr reference at(index_type i) const { return m_p[i]; }
    constexpr Span subspan(index_type offset, index_type count = -1) const
    {
        return Span(m_p + offset, count < 0 ? m_n - offset : count);
    }
    constexpr Span first(index_type n) const { return Span(m_p, n); }
    constexpr Span last(index_type n) const { return Span(m_p + (m_n - n), n); }

    iterator begin() const noexcept { return m_p; }
    iterator end() const noexcept { return m_p + m_n; }
    const_iterator cbegin() const noexcept { return m_p; }
    const_iterator cend() const noexcept { return m_p + m_n; }
    reverse_iterator rbegin() const noexcept { return reverse_iterator(end()); }
    reverse_iterator rend() const noexcept { return reverse_iterator(begin()); }
    const_reverse_iterator crbegin() const noexcept { return const_reverse_iterator(cend()); }
    const_reverse_iterator crend() const noexcept { return const_reverse_iterator(cbegin()); }

private:
    pointer m_p;
    index_type m_n;
};

template <class ElementType>
constexpr Span<ElementType> make_span(ElementType* ptr, typename Span<ElementType>::index_type count)
{
    return Span<ElementType>(ptr, count);
}
template <class ElementType> constexpr Span<ElementType> make_span(ElementType* first, ElementType* last)
{
    return Span<ElementType>(first, last);
}
template <class ElementType, std::size_t N> constexpr Span<ElementType> make_span(ElementType (&arr)[N]) noexcept
{
    return Span<ElementType>(arr);
}
template <class Container, class = std::enable_if_t<!std::is_array<Container>::value>>
constexpr Span<std::remove_pointer_t<decltype(std::declval<Container&>().data())>> make_span(Container& c)
{
    return Span<std::remove_pointer_t<decltype(std::declval<Container&>().data())>>(c.data(),
                                                                                   std::ptrdiff_t(c.size()));
}
template <class Container, class = std::enable_if_t<!std::is_array<Container>::value>>
constexpr Span<std::remove_pointer_t<decltype(std::declval<const Container&>().data())>>
make_span(const Container& c)
{
    return Span<std::remove_pointer_t<decltype(std::declval<const Container&>().data())>>(
        c.data(), std::ptrdiff_t(c.size()));
}

// ------------------------------------------------------------------------------------------------
// Vectors and matrices (row-major)
// ------------------------------------------------------------------------------------------------
template <unsigned int VecSize> class VectorFixSize
{
protected:
    double m_data[VecSize];

public:
    using value_type = double;
    VectorFixSize() = default; // like iDynTree: coefficients are NOT initialised
    VectorFixSize(const double* in, unsigned int n)
    {
        for (unsigned int i = 0; i < VecSize; ++i)
            m_data[i] = i < n ? in[i] : 0.0;
    }
    VectorFixSize(Span<const double> s) : VectorFixSize(s.data(), static_cast<unsigned int>(s.size())) {}

    double operator()(unsigned int i) const
    {
        assert(i < VecSize);
        return m_data[i];
    }
    double& operator()(unsigned int i)
    {
        assert(i < VecSize);
        return m_data[i];
    }
    double operator[](unsigned int i) const { return (*this)(i); }
    double& operator[](unsigned int i) { return (*this)(i); }
    double getVal(unsigned int i) const { return (*this)(i); }
    bool setVal(unsigned int i, double v)
    {
        if (i >= VecSize) return false;
        m_data[i] = v;
        return true;
    }
    unsigned int size() const { return VecSize; }
    const double* data() const { return m_data; }
    double* data() { return m_data; }
    void zero()
    {
        for (unsigned int i = 0; i < VecSize; ++i)
            m_data[i] = 0.0;
    }
    const double* begin() const { return m_data; }
    const double* end() const { return m_data + VecSize; }
    double* begin() { return m_data; }
    double* end() { return m_data + VecSize; }
    std::string toString() const
    {
        std::ostringstream ss;
        for (unsigned int i = 0; i < VecSize; ++i)
            ss << m_data[i] << " ";
        return ss.str();
    }
};
using Vector2 = VectorFixSize<2>;
using Vector3 = VectorFixSize<3>;
using Vector4 = VectorFixSize<4>;
using Vector6 = VectorFixSize<6>;

class VectorDynSize
{
    std::vector<double> m_v;

public:
    using value_type = double;
    VectorDynSize() = default;
    explicit VectorDynSize(std::size_t n) : m_v(n, 0.0) {}
    VectorDynSize(const double* in, std::size_t n) : m_v(in, in + n) {}
    VectorDynSize(Span<const double> s) : m_v(s.begin(), s.end()) {}
    double operator()(std::size_t i) const { return m_v[i]; }
    double& operator()(std::size_t i) { return m_v[i]; }
    double operator[](std::size_t i) const { return m_v[i]; }
    double& operator[](std::size_t i) { return m_v[i]; }
    double getVal(std::size_t i) const { return m_v[i]; }
    bool setVal(std::size_t i, double v)
    {
        if (i >= m_v.size()) return false;
        m_v[i] = v;
        return true;
    }
    std::size_t size() const { return m_v.size(); }
    const double* data() const { return m_v.data(); }
    double* data() { return m_v.data(); }
    void resize(std::size_t n) { m_v.resize(n, 0.0); }
    void reserve(std::size_t n) { m_v.reserve(n); }
    void zero() { std::fill(m_v.begin(), m_v.end(), 0.0); }
    const double* begin() const { return m_v.data(); }
    const double* end() const { return m_v.data() + m_v.size(); }
    double* begin() { return m_v.data(); }
    double* end() { return m_v.data() + m_v.size(); }
    std::string toString() const
    {
        std::ostringstream ss;
        for (double d : m_v)
            ss << d << " ";
        return ss.str();
    }
};

template <unsigned int nRows, unsigned int nCols> class MatrixFixSize
{
protected:
    double m_data[nRows * nCols]; // row-major

public:
    MatrixFixSize() = default; // coefficients NOT initialised, as in iDynTree
    MatrixFixSize(const double* in, unsigned int r, unsigned int c)
    {
        assert(r == nRows && c == nCols);
        (void)r;
        (void)c;
        for (unsigned int i = 0; i < nRows * nCols; ++i)
            m_data[i] = in[i];
    }
    double operator()(unsigned int r, unsigned int c) const
    {
        assert(r < nRows && c < nCols);
        return m_data[r * nCols + c];
    }
    double& operator()(unsigned int r, unsigned int c)
    {
        assert(r < nRows && c < nCols);
        return m_data[r * nCols + c];
    }
    double getVal(unsigned int r, unsigned int c) const { return (*this)(r, c); }
    unsigned int rows() const { return nRows; }
    unsigned int cols() const { return nCols; }
    const double* data() const { return m_data; }
    double* data() { return m_data; }
    void zero()
    {
        for (unsigned int i = 0; i < nRows * nCols; ++i)
            m_data[i] = 0.0;
    }
};
using Matrix3x3 = MatrixFixSize<3, 3>;
using Matrix4x4 = MatrixFixSize<4, 4>;
using Matrix6x6 = MatrixFixSize<6, 6>;

class MatrixDynSize
{
    std::vector<double> m_v; // row-major
    std::size_t m_r = 0, m_c = 0;

public:
    MatrixDynSize() = default;
    MatrixDynSize(std::size_t r, std::size_t c) : m_v(r * c, 0.0), m_r(r), m_c(c) {}
    double operator()(std::size_t r, std::size_t c) const { return m_v[r * m_c + c]; }
    double& operator()(std::size_t r, std::size_t c) { return m_v[r * m_c + c]; }
    double getVal(std::size_t r, std::size_t c) const { return (*this)(r, c); }
    std::size_t rows() const { return m_r; }
    std::size_t cols() const { return m_c; }
    const double* data() const { return m_v.data(); }
    double* data() { return m_v.data(); }
    void resize(std::size_t r, std::size_t c)
    {
        if (r == m_r && c == m_c) return;
        m_v.resize(r * c, 0.0);
        m_r = r;
        m_c = c;
    }
    void zero() { std::fill(m_v.begin(), m_v.end(), 0.0); }
};

// ------------------------------------------------------------------------------------------------
// Geometric types
// ------------------------------------------------------------------------------------------------
class Rotation;

class GeomVector3 : public Vector3
{
public:
    GeomVector3() = default;
    GeomVector3(double x, double y, double z)
    {
        m_data[0] = x;
        m_data[1] = y;
        m_data[2] = z;
    }
    GeomVector3(const double* in, unsigned int n) : Vector3(in, n) {}
};
class Position : public GeomVector3
{
public:
    using GeomVector3::GeomVector3;
    static Position Zero()
    {
        Position p;
        p.zero();
        return p;
    }
};
class LinearMotionVector3 : public GeomVector3
{
public:
    using GeomVector3::GeomVector3;
};
class AngularMotionVector3 : public GeomVector3
{
public:
    using GeomVector3::GeomVector3;
    Rotation exp() const; // Rodrigues' formula
};
class LinearForceVector3 : public GeomVector3
{
public:
    using GeomVector3::GeomVector3;
};
class AngularForceVector3 : public GeomVector3
{
public:
    using GeomVector3::GeomVector3;
};
using Force = LinearForceVector3;
using Torque = AngularForceVector3;
using LinVelocity = LinearMotionVector3;
using AngVelocity = AngularMotionVector3;
using LinAcceleration = LinearMotionVector3;
using AngAcceleration = AngularMotionVector3;

class Rotation : public Matrix3x3
{
public:
    Rotation() = default;
    Rotation(double xx, double xy, double xz, double yx, double yy, double yz, double zx, double zy,
             double zz)
    {
        const double v[9] = {xx, xy, xz, yx, yy, yz, zx, zy, zz};
        for (int i = 0; i < 9; ++i)
            m_data[i] = v[i];
    }
    Rotation(const double* in, unsigned int r, unsigned int c) : Matrix3x3(in, r, c) {}
    static Rotation Identity() { return Rotation(1, 0, 0, 0, 1, 0, 0, 0, 1); }
    static Rotation RotX(double a)
    {
        const double c = std::cos(a), s = std::sin(a);
        return Rotation(1, 0, 0, 0, c, -s, 0, s, c);
    }
    static Rotation RotY(double a)
    {
        const double c = std::cos(a), s = std::sin(a);
        return Rotation(c, 0, s, 0, 1, 0, -s, 0, c);
    }
    static Rotation RotZ(double a)
    {
        const double c = std::cos(a), s = std::sin(a);
        return Rotation(c, -s, 0, s, c, 0, 0, 0, 1);
    }
    static Rotation compose(const Rotation& a, const Rotation& b)
    {
        Rotation out;
        for (unsigned int i = 0; i < 3; ++i)
            for (unsigned int j = 0; j < 3; ++j)
                out(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
        return out;
    }
    /// R = Rz(yaw) * Ry(pitch) * Rx(roll)
    static Rotation RPY(double roll, double pitch, double yaw)
    {
        return compose(RotZ(yaw), compose(RotY(pitch), RotX(roll)));
    }
    Rotation operator*(const Rotation& o) const { return compose(*this, o); }
    Rotation inverse() const
    {
        const Rotation& r = *this;
        return Rotation(r(0, 0), r(1, 0), r(2, 0), r(0, 1), r(1, 1), r(2, 1), r(0, 2), r(1, 2), r(2, 2));
    }
    Position operator*(const Position& p) const
    {
        const Rotation& r = *this;
        return Position(r(0, 0) * p(0) + r(0, 1) * p(1) + r(0, 2) * p(2),
                        r(1, 0) * p(0) + r(1, 1) * p(1) + r(1, 2) * p(2),
                        r(2, 0) * p(0) + r(2, 1) * p(1) + r(2, 2) * p(2));
    }
};

inline Rotation AngularMotionVector3::exp() const
{
    const double x = m_data[0], y = m_data[1], z = m_data[2];
    const double th = std::sqrt(x * x + y * y + z * z);
    double a, b; // R = I + a S + b S^2
    if (th < 1e-10)
    {
        a = 1.0 - th * th / 6.0;
        b = 0.5 - th * th / 24.0;
    } else
    {
        a = std::sin(th) / th;
        b = (1.0 - std::cos(th)) / (th * th);
    }
    const double S[3][3] = {{0, -z, y}, {z, 0, -x}, {-y, x, 0}};
    Rotation out = Rotation::Identity();
    for (unsigned int i = 0; i < 3; ++i)
        for (unsigned int j = 0; j < 3; ++j)
        {
            double s2 = 0;
            for (unsigned int k = 0; k < 3; ++k)
                s2 += S[i][k] * S[k][j];
            out(i, j) += a * S[i][j] + b * s2;
        }
    return out;
}

class Transform
{
    Position pos;
    Rotation rot;

public:
    Transform() = default;
    Transform(const Rotation& r, const Position& p) : pos(p), rot(r) {}
    static Transform Identity() { return Transform(Rotation::Identity(), Position::Zero()); }
    const Position& getPosition() const { return pos; }
    const Rotation& getRotation() const { return rot; }
    void setPosition(const Position& p) { pos = p; }
    void setRotation(const Rotation& r) { rot = r; }
    Transform operator*(const Transform& o) const
    {
        const Position rp = rot * o.pos;
        return Transform(rot * o.rot, Position(rp(0) + pos(0), rp(1) + pos(1), rp(2) + pos(2)));
    }
};

template <class LinT, class AngT> class SpatialVector
{
protected:
    LinT linearVec3;
    AngT angularVec3;

public:
    SpatialVector() = default;
    SpatialVector(const LinT& l, const AngT& a) : linearVec3(l), angularVec3(a) {}
    LinT& getLinearVec3() { return linearVec3; }
    AngT& getAngularVec3() { return angularVec3; }
    const LinT& getLinearVec3() const { return linearVec3; }
    const AngT& getAngularVec3() const { return angularVec3; }
    void setLinearVec3(const LinT& l) { linearVec3 = l; }
    void setAngularVec3(const AngT& a) { angularVec3 = a; }
    double operator()(unsigned int i) const
    {
        assert(i < 6);
        return i < 3 ? linearVec3(i) : angularVec3(i - 3);
    }
    double& operator()(unsigned int i)
    {
        assert(i < 6);
        return i < 3 ? linearVec3(i) : angularVec3(i - 3);
    }
    double operator[](unsigned int i) const { return (*this)(i); }
    double& operator[](unsigned int i) { return (*this)(i); }
    double getVal(unsigned int i) const { return (*this)(i); }
    unsigned int size() const { return 6; }
    void zero()
    {
        linearVec3.zero();
        angularVec3.zero();
    }
    Vector6 asVector() const
    {
        Vector6 v;
        for (unsigned int i = 0; i < 6; ++i)
            v(i) = (*this)(i);
        return v;
    }
};

template <class Derived, class LinT, class AngT> class SpatialVectorZero : public SpatialVector<LinT, AngT>
{
public:
    using SpatialVector<LinT, AngT>::SpatialVector;
    static Derived Zero()
    {
        Derived d;
        d.zero();
        return d;
    }
};
class Twist : public SpatialVectorZero<Twist, LinearMotionVector3, AngularMotionVector3>
{
public:
    using SpatialVectorZero::SpatialVectorZero;
};
class SpatialAcc : public SpatialVectorZero<SpatialAcc, LinearMotionVector3, AngularMotionVector3>
{
public:
    using SpatialVectorZero::SpatialVectorZero;
};
class Wrench : public SpatialVectorZero<Wrench, LinearForceVector3, AngularForceVector3>
{
public:
    using SpatialVectorZero::SpatialVectorZero;
};
using SpatialMotionVector = SpatialVector<LinearMotionVector3, AngularMotionVector3>;
using SpatialForceVector = SpatialVector<LinearForceVector3, AngularForceVector3>;

// ------------------------------------------------------------------------------------------------
// EigenHelpers
// ------------------------------------------------------------------------------------------------
template <unsigned int N> inline Eigen::Map<Eigen::Matrix<double, int(N), 1>> toEigen(VectorFixSize<N>& v)
{
    return Eigen::Map<Eigen::Matrix<double, int(N), 1>>(v.data());
}
template <unsigned int N>
inline Eigen::Map<const Eigen::Matrix<double, int(N), 1>> toEigen(const VectorFixSize<N>& v)
{
    return Eigen::Map<const Eigen::Matrix<double, int(N), 1>>(v.data());
}
inline Eigen::Map<Eigen::VectorXd> toEigen(VectorDynSize& v)
{
    return Eigen::Map<Eigen::VectorXd>(v.data(), Eigen::Index(v.size()));
}
inline Eigen::Map<const Eigen::VectorXd> toEigen(const VectorDynSize& v)
{
    return Eigen::Map<const Eigen::VectorXd>(v.data(), Eigen::Index(v.size()));
}
inline Eigen::Map<Eigen::VectorXd> toEigen(Span<double> s)
{
    return Eigen::Map<Eigen::VectorXd>(s.data(), s.size());
}
template <unsigned int R, unsigned int C>
inline Eigen::Map<Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>> toEigen(MatrixFixSize<R, C>& m)
{
    return Eigen::Map<Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>>(m.data());
}
template <unsigned int R, unsigned int C>
inline Eigen::Map<const Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>>
toEigen(const MatrixFixSize<R, C>& m)
{
    return Eigen::Map<const Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>>(m.data());
}
using EigenDynRowMajor = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;
inline Eigen::Map<EigenDynRowMajor> toEigen(MatrixDynSize& m)
{
    return Eigen::Map<EigenDynRowMajor>(m.data(), Eigen::Index(m.rows()), Eigen::Index(m.cols()));
}
inline Eigen::Map<const EigenDynRowMajor> toEigen(const MatrixDynSize& m)
{
    return Eigen::Map<const EigenDynRowMajor>(m.data(), Eigen::Index(m.rows()), Eigen::Index(m.cols()));
}
// spatial vectors are returned BY VALUE (their two halves are separate members)
template <class LinT, class AngT> inline Eigen::Matrix<double, 6, 1> toEigen(const SpatialVector<LinT, AngT>& s)
{
    Eigen::Matrix<double, 6, 1> out;
    for (unsigned int i = 0; i < 6; ++i)
        out(Eigen::Index(i)) = s(i);
    return out;
}

template <class Derived>
inline Eigen::Matrix<typename Derived::Scalar, 3, 3, Eigen::RowMajor> skew(const Eigen::MatrixBase<Derived>& vec)
{
    assert(vec.size() == 3);
    Eigen::Matrix<typename Derived::Scalar, 3, 3, Eigen::RowMajor> m;
    m(0, 0) = 0.0;
    m(0, 1) = -vec[2];
    m(0, 2) = vec[1];
    m(1, 0) = vec[2];
    m(1, 1) = 0.0;
    m(1, 2) = -vec[0];
    m(2, 0) = -vec[1];
    m(2, 1) = vec[0];
    m(2, 2) = 0.0;
    return m;
}

} // namespace iDynTree

#endif // BLF_REFBUILD_STANDIN_IDYNTREE_CORE
