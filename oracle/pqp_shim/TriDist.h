// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's TriDist.h.
#ifndef PQP_SHIM_TRIDIST_H
#define PQP_SHIM_TRIDIST_H
#include "PQP_Compile.h"
// closest points P on triangle S and Q on triangle T; returns their distance
// (0 when the triangles overlap).
PQP_REAL TriDist(PQP_REAL P[3], PQP_REAL Q[3], const PQP_REAL S[3][3], const PQP_REAL T[3][3]);
// closest points X,Y on segments (P,P+A), (Q,Q+B) and a vector VEC between them.
void SegPoints(PQP_REAL VEC[3], PQP_REAL X[3], PQP_REAL Y[3], const PQP_REAL P[3],
               const PQP_REAL A[3], const PQP_REAL Q[3], const PQP_REAL B[3]);
#endif
