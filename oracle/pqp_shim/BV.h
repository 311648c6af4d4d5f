// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's BV.h
// (fields as used at /root/reference/C2A/src/C2A_BV.cpp:122,160-164,178,342-347;
//  Leaf/GetSize at C2A/src/C2A.cpp:1127,1192).
#ifndef PQP_SHIM_BV_H
#define PQP_SHIM_BV_H
#include <math.h>
#include "Tri.h"
#include "PQP_Compile.h"

struct BV
{
  PQP_REAL R[3][3];   // orientation of RSS & OBB
  PQP_REAL Tr[3];     // position of rectangle
  PQP_REAL l[2];      // side lengths of rectangle
  PQP_REAL r;         // radius of sphere summed with rectangle to form RSS
  PQP_REAL To[3];     // position of obb
  PQP_REAL d[3];      // (half) dimensions of obb
  int first_child;    // >=0: index of first child BV; <0: -(triangle index + 1)

  BV();
  ~BV();              // user-provided on purpose: the reference delete[]s a C2A_BV[] through BV*
  int Leaf() { return first_child < 0; }
  PQP_REAL GetSize() { return (sqrt(l[0] * l[0] + l[1] * l[1]) + 2 * r); }
  void FitToTris(PQP_REAL O[3][3], Tri *tris, int num_tris);
};

int BV_Overlap(PQP_REAL R[3][3], PQP_REAL T[3], BV *b1, BV *b2);
PQP_REAL BV_Distance(PQP_REAL R[3][3], PQP_REAL T[3], BV *b1, BV *b2);
#endif
