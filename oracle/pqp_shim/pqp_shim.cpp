// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP (GammaUNC/PQP, v1.3
// lineage, unpinned upstream: /root/reference/README.md:11), just enough for the
// reference's C2A/src/*.cpp to link.  The geometry on the CCD path (TriDist,
// SegPoints, TriDistance) forwards to the oracle port (oracle/c2a_oracle.cpp),
// which follows the reference's in-tree copies (C2A/src/C2A.cpp:59-424) and is
// checked bit-exactly against them in tests/test_oracle_ref.py.
#include <math.h>
#include <string.h>

#include "PQP.h"
#include "MatVec.h"
#include "TriDist.h"
#include "RectDist.h"
#include "OBB_Disjoint.h"
#include "../c2a_oracle.h"

BV::BV() { first_child = 0; }
BV::~BV() {}

PQP_Model::PQP_Model()
{
  build_state = 0;  // PQP_BUILD_STATE_EMPTY
  tris = 0; num_tris = 0; num_tris_alloced = 0;
  b = 0; num_bvs = 0; num_bvs_alloced = 0;
  last_tri = 0;
}

PQP_Model::~PQP_Model()
{
  if (b != 0) delete[] b;
  if (tris != 0) delete[] tris;
}

PQP_CollideResult::PQP_CollideResult()
{
  pairs = 0; num_pairs = num_pairs_alloced = 0;
  num_bv_tests = 0; num_tri_tests = 0; query_time_secs = 0;
}
PQP_CollideResult::~PQP_CollideResult() { delete[] pairs; }
void PQP_CollideResult::FreePairsList()
{
  delete[] pairs;
  pairs = 0; num_pairs = num_pairs_alloced = 0;
}
void PQP_CollideResult::SizeTo(int n)
{
  if (n < num_pairs) return;
  CollisionPair *np = new CollisionPair[n];
  if (num_pairs > 0) memcpy(np, pairs, sizeof(CollisionPair) * num_pairs);
  delete[] pairs;
  pairs = np;
  num_pairs_alloced = n;
}
void PQP_CollideResult::Add(int a, int b)
{
  if (num_pairs >= num_pairs_alloced) SizeTo(num_pairs_alloced * 2 + 8);
  pairs[num_pairs].id1 = a;
  pairs[num_pairs].id2 = b;
  num_pairs++;
}

void SegPoints(PQP_REAL VEC[3], PQP_REAL X[3], PQP_REAL Y[3], const PQP_REAL P[3], const PQP_REAL A[3],
               const PQP_REAL Q[3], const PQP_REAL B[3])
{
  orc_seg_points(VEC, X, Y, P, A, Q, B);
}

PQP_REAL TriDist(PQP_REAL P[3], PQP_REAL Q[3], const PQP_REAL S[3][3], const PQP_REAL T[3][3])
{
  return orc_tri_dist(P, Q, &S[0][0], &T[0][0]);
}

// call sites: /root/reference/C2A/src/C2A.cpp:1148,1916
PQP_REAL TriDistance(PQP_REAL R[3][3], PQP_REAL T[3], Tri *t1, Tri *t2, PQP_REAL p[3], PQP_REAL q[3])
{
  PQP_REAL tri1[3][3], tri2[3][3];
  VcV(tri1[0], t1->p1); VcV(tri1[1], t1->p2); VcV(tri1[2], t1->p3);
  MxVpV(tri2[0], R, t2->p1, T); MxVpV(tri2[1], R, t2->p2, T); MxVpV(tri2[2], R, t2->p3, T);
  return TriDist(p, q, tri1, tri2);
}

PQP_REAL RectDist(PQP_REAL Rab[3][3], PQP_REAL Tab[3], PQP_REAL a[2], PQP_REAL b[2])
{
  PQP_REAL P[3], Q[3], S[3];
  return orc_rect_dist(&Rab[0][0], Tab, a, b, P, Q, S);
}

// Separating-axis overlap test for two triangles (off the CCD path; used by the
// reference's discrete C2A_Collide only, C2A/src/C2A_PQP.cpp:831).
static int axis_separates(const PQP_REAL ax[3], const PQP_REAL p[3][3], const PQP_REAL q[3][3])
{
  PQP_REAL pmin = VdotV(ax, p[0]), pmax = pmin, qmin = VdotV(ax, q[0]), qmax = qmin;
  for (int i = 1; i < 3; i++)
  {
    PQP_REAL v = VdotV(ax, p[i]); if (v < pmin) pmin = v; if (v > pmax) pmax = v;
    PQP_REAL w = VdotV(ax, q[i]); if (w < qmin) qmin = w; if (w > qmax) qmax = w;
  }
  return (pmin > qmax) || (qmin > pmax);
}

int TriContact(PQP_REAL *P1, PQP_REAL *P2, PQP_REAL *P3, PQP_REAL *Q1, PQP_REAL *Q2, PQP_REAL *Q3)
{
  PQP_REAL p[3][3], q[3][3], e[3][3], f[3][3], n[3], m[3], ax[3];
  VmV(p[0], P1, P1); VmV(p[1], P2, P1); VmV(p[2], P3, P1);
  VmV(q[0], Q1, P1); VmV(q[1], Q2, P1); VmV(q[2], Q3, P1);
  VmV(e[0], p[1], p[0]); VmV(e[1], p[2], p[1]); VmV(e[2], p[0], p[2]);
  VmV(f[0], q[1], q[0]); VmV(f[1], q[2], q[1]); VmV(f[2], q[0], q[2]);
  VcrossV(n, e[0], e[1]);
  VcrossV(m, f[0], f[1]);
  if (axis_separates(n, p, q)) return 0;
  if (axis_separates(m, p, q)) return 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      VcrossV(ax, e[i], f[j]);
      if (axis_separates(ax, p, q)) return 0;
    }
  for (int i = 0; i < 3; i++)
  {
    VcrossV(ax, e[i], n); if (axis_separates(ax, p, q)) return 0;
    VcrossV(ax, f[i], m); if (axis_separates(ax, p, q)) return 0;
  }
  return 1;
}

// 15-axis separating-axis test for oriented boxes (off the CCD path).
int obb_disjoint(PQP_REAL B[3][3], PQP_REAL T[3], PQP_REAL a[3], PQP_REAL b[3])
{
  const PQP_REAL reps = (PQP_REAL)1e-6;
  PQP_REAL Bf[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Bf[i][j] = myfabs(B[i][j]) + reps;
  for (int i = 0; i < 3; i++)
  {
    PQP_REAL t = myfabs(T[i]);
    if (t > a[i] + b[0] * Bf[i][0] + b[1] * Bf[i][1] + b[2] * Bf[i][2]) return 1 + i;
  }
  for (int j = 0; j < 3; j++)
  {
    PQP_REAL s = T[0] * B[0][j] + T[1] * B[1][j] + T[2] * B[2][j];
    PQP_REAL t = myfabs(s);
    if (t > b[j] + a[0] * Bf[0][j] + a[1] * Bf[1][j] + a[2] * Bf[2][j]) return 4 + j;
  }
  int code = 7;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++, code++)
    {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      PQP_REAL s = T[i2] * B[i1][j] - T[i1] * B[i2][j];
      PQP_REAL t = myfabs(s);
      PQP_REAL ra = a[i1] * Bf[i2][j] + a[i2] * Bf[i1][j];
      PQP_REAL rb = b[j1] * Bf[i][j2] + b[j2] * Bf[i][j1];
      if (t > ra + rb) return code;
    }
  return 0;
}

int BV_Overlap(PQP_REAL R[3][3], PQP_REAL T[3], BV *b1, BV *b2)
{
  return (obb_disjoint(R, T, b1->d, b2->d) == 0);
}

PQP_REAL BV_Distance(PQP_REAL R[3][3], PQP_REAL T[3], BV *b1, BV *b2)
{
  PQP_REAL dist = RectDist(R, T, b1->l, b2->l);
  dist -= (b1->r + b2->r);
  return (dist < (PQP_REAL)0.0) ? (PQP_REAL)0.0 : dist;
}
