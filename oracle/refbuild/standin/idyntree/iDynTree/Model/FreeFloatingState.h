// forwards to the stand-in (see StandinModel.h)
#include <iDynTree/Model/StandinModel.h>
