/* forwarding header: lets code written against the reference tree (#include "operators.h") build against this library */
#include "hpgmg_operators.h"
