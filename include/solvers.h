/* forwarding header: lets code written against the reference tree (#include "solvers.h") build against this library */
#include "hpgmg_solvers.h"
