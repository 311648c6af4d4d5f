// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's Tri.h
// (fields as used at /root/reference/C2A/src/C2A_PQP.cpp:164-178).
#ifndef PQP_SHIM_TRI_H
#define PQP_SHIM_TRI_H
#include "PQP_Compile.h"
struct Tri
{
  PQP_REAL p1[3];
  PQP_REAL p2[3];
  PQP_REAL p3[3];
  int id;
};
#endif
