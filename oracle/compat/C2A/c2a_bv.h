// TEST INFRASTRUCTURE (oracle): case alias (/root/reference/C2A/InterpMotion.h:13).
#include "C2A/C2A_BV.h"
