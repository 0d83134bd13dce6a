// Shim: see CoreTypes.h (the real iDynTree is not installed here).
#include <iDynTree/Core/CoreTypes.h>
