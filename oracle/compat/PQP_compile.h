// TEST INFRASTRUCTURE (oracle): case alias (/root/reference/C2A/InterpMotion.h:12).
#include "PQP_Compile.h"
