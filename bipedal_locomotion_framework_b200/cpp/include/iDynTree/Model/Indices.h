// Indices.h -- the index typedefs of iDynTree's interface that the System facade needs
// (FrameIndex identifies the contact frame in ContactWrench).  With the real iDynTree
// (BLF_HAVE_IDYNTREE) include its own header instead.
#ifndef BLF_IDYNTREE_MODEL_INDICES_SHIM_H
#define BLF_IDYNTREE_MODEL_INDICES_SHIM_H

#include <cstddef>

namespace iDynTree
{
typedef std::ptrdiff_t LinkIndex;
typedef std::ptrdiff_t FrameIndex;
constexpr std::ptrdiff_t FRAME_INVALID_INDEX = -1;
} // namespace iDynTree

#endif // BLF_IDYNTREE_MODEL_INDICES_SHIM_H
