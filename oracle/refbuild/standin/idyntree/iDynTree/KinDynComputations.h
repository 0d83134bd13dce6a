// forwards to the stand-in TEST DOUBLE (see iDynTree/Model/StandinModel.h)
#include <iDynTree/Model/StandinModel.h>
