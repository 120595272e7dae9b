/* forwarding header: lets code written against the reference tree (#include "defines.h") build against this library */
#include "hpgmg_defines.h"
