// dyn_kernels_wide.cu -- the 16- and 32-lane size classes of the warp-level mass-matrix solve (32 .. 64
// unknowns), compiled as their own translation unit beside dyn_kernels.cu (same source, other class
// list): see the comment above BLF_LLT_CLASSES there.
#define BLF_LLT_TU_WIDE
#include "dyn_kernels.cu"
