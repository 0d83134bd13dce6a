// Stand-in for iDynTree/Model/Indices.h (test infrastructure, see oracle/refbuild/README.md):
// the index typedefs of iDynTree's published interface.
#ifndef BLF_REFBUILD_STANDIN_IDYNTREE_INDICES
#define BLF_REFBUILD_STANDIN_IDYNTREE_INDICES
#include <cstddef>
namespace iDynTree
{
typedef std::ptrdiff_t LinkIndex;
typedef std::ptrdiff_t JointIndex;
typedef std::ptrdiff_t DOFIndex;
typedef std::ptrdiff_t FrameIndex;
typedef std::ptrdiff_t TraversalIndex;
constexpr std::ptrdiff_t LINK_INVALID_INDEX = -1;
constexpr std::ptrdiff_t JOINT_INVALID_INDEX = -1;
constexpr std::ptrdiff_t DOF_INVALID_INDEX = -1;
constexpr std::ptrdiff_t FRAME_INVALID_INDEX = -1;
} // namespace iDynTree
#endif
