// c2a_b200 drop-in: the reference splits its public API over several headers (C2A/C2A_Internal.h);
// here everything lives in C2A/C2A.h.
#include "C2A.h"
