// TEST INFRASTRUCTURE (oracle): case alias -- the reference was written for a
// case-insensitive file system (/root/reference/C2A/src/InterpMotion.cpp:1).
#include "C2A/InterpMotion.h"
