// EigenHelpers for THE PRODUCT FACADE'S iDynTree-layout types
// (bipedal_locomotion_framework_b200/cpp/include/iDynTree/Core/CoreTypes.h), over the stand-in Eigen.
//
// TEST INFRASTRUCTURE ONLY.  Purpose: compile the reference's own, UNMODIFIED Catch2 test
// (src/ContactModels/tests/ContinousContactModelTest.cpp) against the B200 facade instead of the
// reference's classes -- the drop-in check: same test source, other implementation behind the same
// interface.  Include path when building it: the facade's include dir, this directory, and
// oracle/refbuild/standin/eigen + standin/catch (NOT standin/idyntree: the iDynTree-layout types
// are the facade's own).
#ifndef BLF_REFBUILD_FACADE_GLUE_EIGEN_HELPERS
#define BLF_REFBUILD_FACADE_GLUE_EIGEN_HELPERS

#include <Eigen/Core>
#include <iDynTree/Core/CoreTypes.h>

namespace iDynTree
{
template <unsigned N> inline Eigen::Map<Eigen::Matrix<double, int(N), 1>> toEigen(VectorFixSize<N>& v)
{
    return Eigen::Map<Eigen::Matrix<double, int(N), 1>>(v.data());
}
template <unsigned N> inline Eigen::Map<const Eigen::Matrix<double, int(N), 1>> toEigen(const VectorFixSize<N>& v)
{
    return Eigen::Map<const Eigen::Matrix<double, int(N), 1>>(v.data());
}
template <unsigned R, unsigned C>
inline Eigen::Map<Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>> toEigen(MatrixFixSize<R, C>& m)
{
    return Eigen::Map<Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>>(m.data());
}
template <unsigned R, unsigned C>
inline Eigen::Map<const Eigen::Matrix<double, int(R), int(C), Eigen::RowMajor>> toEigen(const MatrixFixSize<R, C>& m)
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
inline Eigen::Map<Eigen::VectorXd> toEigen(VectorDynSize& v)
{
    return Eigen::Map<Eigen::VectorXd>(v.data(), Eigen::Index(v.size()));
}
inline Eigen::Map<const Eigen::VectorXd> toEigen(const VectorDynSize& v)
{
    return Eigen::Map<const Eigen::VectorXd>(v.data(), Eigen::Index(v.size()));
}
} // namespace iDynTree

#endif
