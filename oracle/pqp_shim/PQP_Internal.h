// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's PQP_Internal.h
// (members as used at /root/reference/C2A/src/C2A_PQP.cpp:87-98,384-401,927-967,
//  1009-1053,1776-1790).  BUILD_STATE is deliberately NOT defined here: the
//  reference re-declares it (C2A/src/C2A.cpp:52-57).
#ifndef PQP_SHIM_INTERNAL_H
#define PQP_SHIM_INTERNAL_H
#include "Tri.h"
#include "BV.h"

class PQP_Model
{
public:
  int build_state;
  Tri *tris;
  int num_tris;
  int num_tris_alloced;
  BV *b;
  int num_bvs;
  int num_bvs_alloced;
  Tri *last_tri;   // closest tri on this model in last distance test

  BV *child(int n) { return &b[n]; }

  PQP_Model();
  ~PQP_Model();
};

struct CollisionPair
{
  int id1;
  int id2;
};

struct PQP_CollideResult
{
  int num_bv_tests;
  int num_tri_tests;
  double query_time_secs;
  PQP_REAL R[3][3];
  PQP_REAL T[3];
  int num_pairs_alloced;
  int num_pairs;
  CollisionPair *pairs;

  void SizeTo(int n);
  void Add(int i1, int i2);
  PQP_CollideResult();
  ~PQP_CollideResult();
  int NumBVTests() { return num_bv_tests; }
  int NumTriTests() { return num_tri_tests; }
  double QueryTimeSecs() { return query_time_secs; }
  void FreePairsList();
  int Colliding() { return (num_pairs > 0); }
  int NumPairs() { return num_pairs; }
  int Id1(int k) { return pairs[k].id1; }
  int Id2(int k) { return pairs[k].id2; }
};

struct PQP_DistanceResult
{
  int num_bv_tests;
  int num_tri_tests;
  double query_time_secs;
  PQP_REAL R[3][3];
  PQP_REAL T[3];
  PQP_REAL rel_err;
  PQP_REAL abs_err;
  PQP_REAL distance;
  PQP_REAL p1[3];
  PQP_REAL p2[3];
  int qsize;

  int NumBVTests() { return num_bv_tests; }
  int NumTriTests() { return num_tri_tests; }
  double QueryTimeSecs() { return query_time_secs; }
  PQP_REAL Distance() { return distance; }
  const PQP_REAL *P1() { return p1; }
  const PQP_REAL *P2() { return p2; }
};

struct PQP_ToleranceResult
{
  int num_bv_tests;
  int num_tri_tests;
  double query_time_secs;
  PQP_REAL R[3][3];
  PQP_REAL T[3];
  int closer_than_tolerance;
  PQP_REAL tolerance;
  PQP_REAL distance;
  PQP_REAL p1[3];
  PQP_REAL p2[3];
  int qsize;

  int NumBVTests() { return num_bv_tests; }
  int NumTriTests() { return num_tri_tests; }
  double QueryTimeSecs() { return query_time_secs; }
  PQP_REAL Distance() { return distance; }
  const PQP_REAL *P1() { return p1; }
  const PQP_REAL *P2() { return p2; }
  int CloserThanTolerance() { return closer_than_tolerance; }
};
#endif
