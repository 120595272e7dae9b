/* forwarding header: lets code written against the reference tree (#include "mg.h") build against this library */
#include "hpgmg_mg.h"
