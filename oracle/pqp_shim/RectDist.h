// TEST INFRASTRUCTURE (oracle) -- PQP's RectDist.h entry point.  Only named by a
// branch that is compiled out when OBB_TYPE is on (C2A/src/C2A_BV.cpp:645-649).
#ifndef PQP_SHIM_RECTDIST_H
#define PQP_SHIM_RECTDIST_H
#include "PQP_Compile.h"
PQP_REAL RectDist(PQP_REAL Rab[3][3], PQP_REAL Tab[3], PQP_REAL a[2], PQP_REAL b[2]);
#endif
