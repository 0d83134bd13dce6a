// Model.h -- the two things FloatingBaseDynamicalSystem asks of iDynTree::Model
// (src/System/src/FloatingBaseSystemDynamics.cpp:64, :205-207): the number of internal degrees of
// freedom and, for its error messages, the link a frame is attached to.  iDynTree (third party, not
// vendored) is absent from this build: this stand-in carries those two answers and nothing of the
// robot description.  With the real iDynTree (BLF_HAVE_IDYNTREE) include its own header instead.
#ifndef BLF_IDYNTREE_MODEL_MODEL_SHIM_H
#define BLF_IDYNTREE_MODEL_MODEL_SHIM_H

#include <cstddef>

#include <iDynTree/Model/Indices.h>

namespace iDynTree
{

class Model
{
    std::size_t m_internalDoFs{0};

public:
    Model() = default;
    explicit Model(std::size_t internalDoFs) : m_internalDoFs(internalDoFs) {}
    std::size_t getNrOfDOFs() const { return m_internalDoFs; }
    /** Frames of this stand-in are attached to the link of the same index. */
    LinkIndex getFrameLink(FrameIndex frame) const { return frame; }
};

} // namespace iDynTree

#endif // BLF_IDYNTREE_MODEL_MODEL_SHIM_H
