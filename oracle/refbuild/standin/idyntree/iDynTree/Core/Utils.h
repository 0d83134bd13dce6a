// forwards to the single stand-in header (see StandinCore.h)
#include <iDynTree/Core/StandinCore.h>
