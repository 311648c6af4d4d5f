// TEST INFRASTRUCTURE (oracle) -- PQP's OBB_Disjoint.h entry point (off the CCD
// path; called by C2A_BV_Overlap, C2A/src/C2A_BV.cpp:644).
#ifndef PQP_SHIM_OBB_DISJOINT_H
#define PQP_SHIM_OBB_DISJOINT_H
#include "PQP_Compile.h"
// B is the rotation and T the translation taking box b's frame to box a's;
// a, b are half-dimensions.  Returns 0 iff the boxes overlap.
int obb_disjoint(PQP_REAL B[3][3], PQP_REAL T[3], PQP_REAL a[3], PQP_REAL b[3]);
#endif
