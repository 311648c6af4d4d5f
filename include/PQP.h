// c2a_b200 drop-in: stands in for the PQP header the reference includes (PQP_REAL, Tri, PQP_OK ...).
#include "C2A/C2A.h"
