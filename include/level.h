/* forwarding header: lets code written against the reference tree (#include "level.h") build against this library */
#include "hpgmg_level.h"
