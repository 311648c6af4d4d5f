// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.  See c2a_oracle.h.
//
// Scalar FP64 restatement of the reference's controlled-CA CCD path.  Paths are
// relative to /root/reference.  Arithmetic order follows the reference
// expression by expression (sums left to right, no FMA: build with
// -ffp-contract=off) because the traversal's predicates are bit-sensitive.
#include "c2a_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
#include <algorithm>
#include <limits>

namespace {

// ---- 3-vector / 3x3 helpers (conventions of PQP MatVec.h as implied by the
// reference's call sites, C2A/src/C2A_PQP.cpp:296-299,934-941) ------------------
inline void v_sub(double r[3], const double a[3], const double b[3]) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
inline void v_add(double r[3], const double a[3], const double b[3]) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
inline void v_cpy(double r[3], const double a[3]) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
inline void v_madd(double r[3], const double a[3], const double b[3], double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
inline double v_dot(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]); }
inline void v_cross(double r[3], const double a[3], const double b[3])
{
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
inline double v_dist2(const double a[3], const double b[3])
{
  return ((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}
inline double v_len(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline void v_normalize(double a[3])
{
  double d = 1.0 / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  a[0] *= d; a[1] *= d; a[2] *= d;
}
// r = M v
inline void m_v(double r[3], const double M[9], const double v[3])
{
  r[0] = (M[0] * v[0] + M[1] * v[1] + M[2] * v[2]);
  r[1] = (M[3] * v[0] + M[4] * v[1] + M[5] * v[2]);
  r[2] = (M[6] * v[0] + M[7] * v[1] + M[8] * v[2]);
}
// r = M v + t
inline void m_v_p(double r[3], const double M[9], const double v[3], const double t[3])
{
  r[0] = (M[0] * v[0] + M[1] * v[1] + M[2] * v[2] + t[0]);
  r[1] = (M[3] * v[0] + M[4] * v[1] + M[5] * v[2] + t[1]);
  r[2] = (M[6] * v[0] + M[7] * v[1] + M[8] * v[2] + t[2]);
}
// r = M^T v
inline void mt_v(double r[3], const double M[9], const double v[3])
{
  r[0] = (M[0] * v[0] + M[3] * v[1] + M[6] * v[2]);
  r[1] = (M[1] * v[0] + M[4] * v[1] + M[7] * v[2]);
  r[2] = (M[2] * v[0] + M[5] * v[1] + M[8] * v[2]);
}
// r = A B
inline void m_m(double r[9], const double A[9], const double B[9])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r[3 * i + j] = (A[3 * i + 0] * B[0 + j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]);
}
// r = A^T B
inline void mt_m(double r[9], const double A[9], const double B[9])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      r[3 * i + j] = (A[0 + i] * B[0 + j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j]);
}

// ---- rectangle-rectangle distance ------------------------------------------
// C2A/C2A_RectDist.h:44-50
inline void clamp_to(double &v, double lo, double hi) { if (v < lo) v = lo; else if (v > hi) v = hi; }

// C2A/C2A_RectDist.h:81-111 (Lumelsky segment-segment parameters)
inline void seg_params(double &t, double &u, double a, double b, double AdB, double AdT, double BdT)
{
  double denom = 1 - (AdB) * (AdB);
  if (denom == 0) t = 0;
  else { t = (AdT - BdT * AdB) / denom; clamp_to(t, 0, a); }
  u = t * AdB - BdT;
  if (u < 0) { u = 0; t = AdT; clamp_to(t, 0, a); }
  else if (u > b) { u = b; t = u * AdB + AdT; clamp_to(t, 0, a); }
}

// C2A/C2A_RectDist.h:123-154
inline bool in_voronoi(double a, double b, double AnB, double AnT, double AdB, double AdT, double BdT)
{
  if (((AnB < 0) ? -AnB : AnB) < 1e-7) return false;
  double t, u, v;
  u = -AnT / AnB; clamp_to(u, 0, b);
  t = u * AdB + AdT; clamp_to(t, 0, a);
  v = t * AdB - BdT;
  if (AnB > 0) { if (v > (u + 1e-7)) return true; }
  else { if (v < (u - 1e-7)) return true; }
  return false;
}

struct RectCtx
{
  const double *R, *T, *a, *b;
  double *P, *Q, *S;
};

// Closest points once an edge pair (A-edge along axis ea on the upper/lower side
// ua; B-edge along axis eb, side ub) is accepted: the P/Q/S block that follows
// each ClosestPoint call in C2A/C2A_RectDist.h:233-848.
inline double rect_edge_finish(const RectCtx &c, int ea, bool ua, int eb, bool ub, double t, double u)
{
  const int oa = 1 - ea, ob = 1 - eb;
  c.P[ea] = t; c.P[oa] = ua ? c.a[oa] : 0; c.P[2] = 0;
  for (int k = 0; k < 3; k++)
  {
    if (ub) c.Q[k] = c.T[k] + c.R[3 * k + ob] * c.b[ob] + c.R[3 * k + eb] * u;
    else c.Q[k] = c.T[k] + c.R[3 * k + eb] * u;
  }
  c.S[0] = c.Q[0] - c.P[0]; c.S[1] = c.Q[1] - c.P[1]; c.S[2] = c.Q[2] - c.P[2];
  return sqrt(v_dot(c.S, c.S));
}

}  // namespace

// C2A/C2A_RectDist.h:157-934 (C2ARectDist).  S = Q - P is left untouched when
// both face separations are negative (the reference does the same).
extern "C" double orc_rect_dist(const double Rab[9], const double Tab[3], const double a[2],
                                const double b[2], double P[3], double Q[3], double S[3])
{
  RectCtx cx = {Rab, Tab, a, b, P, Q, S};
  const double A0B0 = Rab[0], A0B1 = Rab[1], A1B0 = Rab[3], A1B1 = Rab[4];
  const double aA0B0 = a[0] * A0B0, aA0B1 = a[0] * A0B1, aA1B0 = a[1] * A1B0, aA1B1 = a[1] * A1B1;
  const double bA0B0 = b[0] * A0B0, bA1B0 = b[0] * A1B0, bA0B1 = b[1] * A0B1, bA1B1 = b[1] * A1B1;
  double Tba[3];
  mt_v(Tba, Rab, Tab);
  double t, u;

#define RD_EDGE(c1, c2, trivA, ivA, trivB, ivB, ea, ua, eb, ub, la, lb, AdB, AdT, BdT) \
  if ((c1) && (c2))                                                                    \
  {                                                                                    \
    if (((trivA) || in_voronoi ivA) && ((trivB) || in_voronoi ivB))                    \
    {                                                                                  \
      seg_params(t, u, la, lb, AdB, AdT, BdT);                                         \
      return rect_edge_finish(cx, ea, ua, eb, ub, t, u);                               \
    }                                                                                  \
  }

  // --- A1 edges against B1 edges: extents along B0 (x in B) and A0 (x in A) -- :193-360
  double ALL_x = -Tba[0], ALU_x = ALL_x + aA1B0, AUL_x = ALL_x + aA0B0, AUU_x = ALU_x + aA0B0;
  double LA1_lx, LA1_ux, UA1_lx, UA1_ux;
  if (ALL_x < ALU_x) { LA1_lx = ALL_x; LA1_ux = ALU_x; UA1_lx = AUL_x; UA1_ux = AUU_x; }
  else { LA1_lx = ALU_x; LA1_ux = ALL_x; UA1_lx = AUU_x; UA1_ux = AUL_x; }
  double BLL_x = Tab[0], BLU_x = BLL_x + bA0B1, BUL_x = BLL_x + bA0B0, BUU_x = BLU_x + bA0B0;
  double LB1_lx, LB1_ux, UB1_lx, UB1_ux;
  if (BLL_x < BLU_x) { LB1_lx = BLL_x; LB1_ux = BLU_x; UB1_lx = BUL_x; UB1_ux = BUU_x; }
  else { LB1_lx = BLU_x; LB1_ux = BLL_x; UB1_lx = BUU_x; UB1_ux = BUL_x; }

  RD_EDGE(UA1_ux > b[0], UB1_ux > a[0],
          UA1_lx > b[0], (b[1], a[1], A1B0, aA0B0 - b[0] - Tba[0], A1B1, aA0B1 - Tba[1], -Tab[1] - bA1B0),
          UB1_lx > a[0], (a[1], b[1], A0B1, Tab[0] + bA0B0 - a[0], A1B1, Tab[1] + bA1B0, Tba[1] - aA0B1),
          1, true, 1, true, a[1], b[1], A1B1, Tab[1] + bA1B0, Tba[1] - aA0B1)
  RD_EDGE(UA1_lx < 0, LB1_ux > a[0],
          UA1_ux < 0, (b[1], a[1], -A1B0, Tba[0] - aA0B0, A1B1, aA0B1 - Tba[1], -Tab[1]),
          LB1_lx > a[0], (a[1], b[1], A0B1, Tab[0] - a[0], A1B1, Tab[1], Tba[1] - aA0B1),
          1, true, 1, false, a[1], b[1], A1B1, Tab[1], Tba[1] - aA0B1)
  RD_EDGE(LA1_ux > b[0], UB1_lx < 0,
          LA1_lx > b[0], (b[1], a[1], A1B0, -Tba[0] - b[0], A1B1, -Tba[1], -Tab[1] - bA1B0),
          UB1_ux < 0, (a[1], b[1], -A0B1, -Tab[0] - bA0B0, A1B1, Tab[1] + bA1B0, Tba[1]),
          1, false, 1, true, a[1], b[1], A1B1, Tab[1] + bA1B0, Tba[1])
  RD_EDGE(LA1_lx < 0, LB1_lx < 0,
          LA1_ux < 0, (b[1], a[1], -A1B0, Tba[0], A1B1, -Tba[1], -Tab[1]),
          LB1_ux < 0, (a[1], b[1], -A0B1, -Tab[0], A1B1, Tab[1], Tba[1]),
          1, false, 1, false, a[1], b[1], A1B1, Tab[1], Tba[1])

  // --- A1 edges against B0 edges: extents along B1 (y in B) and A0 -- :362-543
  double ALL_y = -Tba[1], ALU_y = ALL_y + aA1B1, AUL_y = ALL_y + aA0B1, AUU_y = ALU_y + aA0B1;
  double LA1_ly, LA1_uy, UA1_ly, UA1_uy;
  if (ALL_y < ALU_y) { LA1_ly = ALL_y; LA1_uy = ALU_y; UA1_ly = AUL_y; UA1_uy = AUU_y; }
  else { LA1_ly = ALU_y; LA1_uy = ALL_y; UA1_ly = AUU_y; UA1_uy = AUL_y; }
  double LB0_lx, LB0_ux, UB0_lx, UB0_ux;
  if (BLL_x < BUL_x) { LB0_lx = BLL_x; LB0_ux = BUL_x; UB0_lx = BLU_x; UB0_ux = BUU_x; }
  else { LB0_lx = BUL_x; LB0_ux = BLL_x; UB0_lx = BUU_x; UB0_ux = BLU_x; }

  RD_EDGE(UA1_uy > b[1], UB0_ux > a[0],
          UA1_ly > b[1], (b[0], a[1], A1B1, aA0B1 - Tba[1] - b[1], A1B0, aA0B0 - Tba[0], -Tab[1] - bA1B1),
          UB0_lx > a[0], (a[1], b[0], A0B0, Tab[0] - a[0] + bA0B1, A1B0, Tab[1] + bA1B1, Tba[0] - aA0B0),
          1, true, 0, true, a[1], b[0], A1B0, Tab[1] + bA1B1, Tba[0] - aA0B0)
  RD_EDGE(UA1_ly < 0, LB0_ux > a[0],
          UA1_uy < 0, (b[0], a[1], -A1B1, Tba[1] - aA0B1, A1B0, aA0B0 - Tba[0], -Tab[1]),
          LB0_lx > a[0], (a[1], b[0], A0B0, Tab[0] - a[0], A1B0, Tab[1], Tba[0] - aA0B0),
          1, true, 0, false, a[1], b[0], A1B0, Tab[1], Tba[0] - aA0B0)
  RD_EDGE(LA1_uy > b[1], UB0_lx < 0,
          LA1_ly > b[1], (b[0], a[1], A1B1, -Tba[1] - b[1], A1B0, -Tba[0], -Tab[1] - bA1B1),
          UB0_ux < 0, (a[1], b[0], -A0B0, -Tab[0] - bA0B1, A1B0, Tab[1] + bA1B1, Tba[0]),
          1, false, 0, true, a[1], b[0], A1B0, Tab[1] + bA1B1, Tba[0])
  RD_EDGE(LA1_ly < 0, LB0_lx < 0,
          LA1_uy < 0, (b[0], a[1], -A1B1, Tba[1], A1B0, -Tba[0], -Tab[1]),
          LB0_ux < 0, (a[1], b[0], -A0B0, -Tab[0], A1B0, Tab[1], Tba[0]),
          1, false, 0, false, a[1], b[0], A1B0, Tab[1], Tba[0])

  // --- A0 edges against B1 edges: extents along B0 and A1 (y in A) -- :545-725
  double BLL_y = Tab[1], BLU_y = BLL_y + bA1B1, BUL_y = BLL_y + bA1B0, BUU_y = BLU_y + bA1B0;
  double LA0_lx, LA0_ux, UA0_lx, UA0_ux;
  if (ALL_x < AUL_x) { LA0_lx = ALL_x; LA0_ux = AUL_x; UA0_lx = ALU_x; UA0_ux = AUU_x; }
  else { LA0_lx = AUL_x; LA0_ux = ALL_x; UA0_lx = AUU_x; UA0_ux = ALU_x; }
  double LB1_ly, LB1_uy, UB1_ly, UB1_uy;
  if (BLL_y < BLU_y) { LB1_ly = BLL_y; LB1_uy = BLU_y; UB1_ly = BUL_y; UB1_uy = BUU_y; }
  else { LB1_ly = BLU_y; LB1_uy = BLL_y; UB1_ly = BUU_y; UB1_uy = BUL_y; }

  RD_EDGE(UA0_ux > b[0], UB1_uy > a[1],
          UA0_lx > b[0], (b[1], a[0], A0B0, aA1B0 - Tba[0] - b[0], A0B1, aA1B1 - Tba[1], -Tab[0] - bA0B0),
          UB1_ly > a[1], (a[0], b[1], A1B1, Tab[1] - a[1] + bA1B0, A0B1, Tab[0] + bA0B0, Tba[1] - aA1B1),
          0, true, 1, true, a[0], b[1], A0B1, Tab[0] + bA0B0, Tba[1] - aA1B1)
  RD_EDGE(UA0_lx < 0, LB1_uy > a[1],
          UA0_ux < 0, (b[1], a[0], -A0B0, Tba[0] - aA1B0, A0B1, aA1B1 - Tba[1], -Tab[0]),
          LB1_ly > a[1], (a[0], b[1], A1B1, Tab[1] - a[1], A0B1, Tab[0], Tba[1] - aA1B1),
          0, true, 1, false, a[0], b[1], A0B1, Tab[0], Tba[1] - aA1B1)
  RD_EDGE(LA0_ux > b[0], UB1_ly < 0,
          LA0_lx > b[0], (b[1], a[0], A0B0, -b[0] - Tba[0], A0B1, -Tba[1], -bA0B0 - Tab[0]),
          UB1_uy < 0, (a[0], b[1], -A1B1, -Tab[1] - bA1B0, A0B1, Tab[0] + bA0B0, Tba[1]),
          0, false, 1, true, a[0], b[1], A0B1, Tab[0] + bA0B0, Tba[1])
  RD_EDGE(LA0_lx < 0, LB1_ly < 0,
          LA0_ux < 0, (b[1], a[0], -A0B0, Tba[0], A0B1, -Tba[1], -Tab[0]),
          LB1_uy < 0, (a[0], b[1], -A1B1, -Tab[1], A0B1, Tab[0], Tba[1]),
          0, false, 1, false, a[0], b[1], A0B1, Tab[0], Tba[1])

  // --- A0 edges against B0 edges: extents along B1 and A1 -- :727-848
  double LA0_ly, LA0_uy, UA0_ly, UA0_uy;
  if (ALL_y < AUL_y) { LA0_ly = ALL_y; LA0_uy = AUL_y; UA0_ly = ALU_y; UA0_uy = AUU_y; }
  else { LA0_ly = AUL_y; LA0_uy = ALL_y; UA0_ly = AUU_y; UA0_uy = ALU_y; }
  double LB0_ly, LB0_uy, UB0_ly, UB0_uy;
  if (BLL_y < BUL_y) { LB0_ly = BLL_y; LB0_uy = BUL_y; UB0_ly = BLU_y; UB0_uy = BUU_y; }
  else { LB0_ly = BUL_y; LB0_uy = BLL_y; UB0_ly = BUU_y; UB0_uy = BLU_y; }

  RD_EDGE(UA0_uy > b[1], UB0_uy > a[1],
          UA0_ly > b[1], (b[0], a[0], A0B1, aA1B1 - Tba[1] - b[1], A0B0, aA1B0 - Tba[0], -Tab[0] - bA0B1),
          UB0_ly > a[1], (a[0], b[0], A1B0, Tab[1] - a[1] + bA1B1, A0B0, Tab[0] + bA0B1, Tba[0] - aA1B0),
          0, true, 0, true, a[0], b[0], A0B0, Tab[0] + bA0B1, Tba[0] - aA1B0)
  RD_EDGE(UA0_ly < 0, LB0_uy > a[1],
          UA0_uy < 0, (b[0], a[0], -A0B1, Tba[1] - aA1B1, A0B0, aA1B0 - Tba[0], -Tab[0]),
          LB0_ly > a[1], (a[0], b[0], A1B0, Tab[1] - a[1], A0B0, Tab[0], Tba[0] - aA1B0),
          0, true, 0, false, a[0], b[0], A0B0, Tab[0], Tba[0] - aA1B0)
  RD_EDGE(LA0_uy > b[1], UB0_ly < 0,
          LA0_ly > b[1], (b[0], a[0], A0B1, -Tba[1] - b[1], A0B0, -Tba[0], -Tab[0] - bA0B1),
          UB0_uy < 0, (a[0], b[0], -A1B0, -Tab[1] - bA1B1, A0B0, Tab[0] + bA0B1, Tba[0]),
          0, false, 0, true, a[0], b[0], A0B0, Tab[0] + bA0B1, Tba[0])
  RD_EDGE(LA0_ly < 0, LB0_ly < 0,
          LA0_uy < 0, (b[0], a[0], -A0B1, Tba[1], A0B0, -Tba[0], -Tab[0]),
          LB0_uy < 0, (a[0], b[0], -A1B0, -Tab[1], A0B0, Tab[0], Tba[0]),
          0, false, 0, false, a[0], b[0], A0B0, Tab[0], Tba[0])
#undef RD_EDGE

  // --- no edge pair: separation along the two face normals -- :850-933
  double sep1, sep2;
  if (Tab[2] > 0.0)
  {
    sep1 = Tab[2];
    if (Rab[6] < 0.0) sep1 += b[0] * Rab[6];
    if (Rab[7] < 0.0) sep1 += b[1] * Rab[7];
  }
  else
  {
    sep1 = -Tab[2];
    if (Rab[6] > 0.0) sep1 -= b[0] * Rab[6];
    if (Rab[7] > 0.0) sep1 -= b[1] * Rab[7];
  }
  if (Tba[2] < 0)
  {
    sep2 = -Tba[2];
    if (Rab[2] < 0.0) sep2 += a[0] * Rab[2];
    if (Rab[5] < 0.0) sep2 += a[1] * Rab[5];
  }
  else
  {
    sep2 = Tba[2];
    if (Rab[2] > 0.0) sep2 -= a[0] * Rab[2];
    if (Rab[5] > 0.0) sep2 -= a[1] * Rab[5];
  }
  if (sep1 >= sep2 && sep1 >= 0)
  {
    P[0] = P[1] = P[2] = 0.0;
    Q[0] = Q[1] = 0.0;
    Q[2] = (Tab[2] > 0.0) ? sep1 : -sep1;
    S[0] = Q[0] - P[0]; S[1] = Q[1] - P[1]; S[2] = Q[2] - P[2];
  }
  if (sep2 >= sep1 && sep2 >= 0)
  {
    Q[0] = Tab[0]; Q[1] = Tab[1]; Q[2] = Tab[2];
    if (Tba[2] < 0)
    {
      P[0] = Rab[2] * sep2 + Tab[0]; P[1] = Rab[5] * sep2 + Tab[1]; P[2] = Rab[8] * sep2 + Tab[2];
    }
    else
    {
      P[0] = -Rab[2] * sep2 + Tab[0]; P[1] = -Rab[5] * sep2 + Tab[1]; P[2] = -Rab[8] * sep2 + Tab[2];
    }
    S[0] = Q[0] - P[0]; S[1] = Q[1] - P[1]; S[2] = Q[2] - P[2];
  }
  double sep = (sep1 > sep2 ? sep1 : sep2);
  return (sep > 0 ? sep : 0);
}

// ---- triangle-triangle distance --------------------------------------------
// PQP SegPoints, as the reference's in-tree copy "ClosestPoints", C2A/src/C2A.cpp:59-163.
extern "C" void orc_seg_points(double VEC[3], double X[3], double Y[3], const double P[3],
                               const double A[3], const double Q[3], const double B[3])
{
  double T[3], TMP[3];
  v_sub(T, Q, P);
  const double AdA = v_dot(A, A), BdB = v_dot(B, B), AdB = v_dot(A, B);
  const double AdT = v_dot(A, T), BdT = v_dot(B, T);
  double denom = AdA * BdB - AdB * AdB;
  double t = (AdT * BdB - BdT * AdB) / denom;
  if ((t < 0) || isnan(t)) t = 0; else if (t > 1) t = 1;
  double u = (t * AdB - BdT) / BdB;

  if ((u <= 0) || isnan(u))
  {
    v_cpy(Y, Q);
    t = AdT / AdA;
    if ((t <= 0) || isnan(t)) { v_cpy(X, P); v_sub(VEC, Q, P); }
    else if (t >= 1) { v_add(X, P, A); v_sub(VEC, Q, X); }
    else { v_madd(X, P, A, t); v_cross(TMP, T, A); v_cross(VEC, A, TMP); }
  }
  else if (u >= 1)
  {
    v_add(Y, Q, B);
    t = (AdB + AdT) / AdA;
    if ((t <= 0) || isnan(t)) { v_cpy(X, P); v_sub(VEC, Y, P); }
    else if (t >= 1) { v_add(X, P, A); v_sub(VEC, Y, X); }
    else { v_madd(X, P, A, t); v_sub(T, Y, P); v_cross(TMP, T, A); v_cross(VEC, A, TMP); }
  }
  else
  {
    v_madd(Y, Q, B, u);
    if ((t <= 0) || isnan(t)) { v_cpy(X, P); v_cross(TMP, T, B); v_cross(VEC, B, TMP); }
    else if (t >= 1) { v_add(X, P, A); v_sub(T, Q, X); v_cross(TMP, T, B); v_cross(VEC, B, TMP); }
    else
    {
      v_madd(X, P, A, t);
      v_cross(VEC, A, B);
      if (v_dot(VEC, T) < 0) { VEC[0] = VEC[0] * -1; VEC[1] = VEC[1] * -1; VEC[2] = VEC[2] * -1; }
    }
  }
}

namespace {
// vertex-of-T against face-of-S test shared by both orientations, C2A/src/C2A.cpp:282-339 / :346-391.
// Returns true and fills (onFace, vertex) when the closest pair is (face of F, vertex of G).
inline bool face_vertex_case(const double F[9], const double Fv[9], const double G[9],
                             int &shown_disjoint, double onFace[3], double vertex[3])
{
  double n[3], V[3], Z[3];
  v_cross(n, &Fv[0], &Fv[3]);
  double nl = v_dot(n, n);
  if (!(nl > 1e-15)) return false;
  double proj[3];
  v_sub(V, &F[0], &G[0]); proj[0] = v_dot(V, n);
  v_sub(V, &F[0], &G[3]); proj[1] = v_dot(V, n);
  v_sub(V, &F[0], &G[6]); proj[2] = v_dot(V, n);
  int point = -1;
  if ((proj[0] > 0) && (proj[1] > 0) && (proj[2] > 0))
  {
    if (proj[0] < proj[1]) point = 0; else point = 1;
    if (proj[2] < proj[point]) point = 2;
  }
  else if ((proj[0] < 0) && (proj[1] < 0) && (proj[2] < 0))
  {
    if (proj[0] > proj[1]) point = 0; else point = 1;
    if (proj[2] > proj[point]) point = 2;
  }
  if (point < 0) return false;
  shown_disjoint = 1;
  const double *g = &G[3 * point];
  for (int e = 0; e < 3; e++)
  {
    v_sub(V, g, &F[3 * e]);
    v_cross(Z, n, &Fv[3 * e]);
    if (!(v_dot(V, Z) > 0)) return false;
  }
  v_madd(onFace, g, n, proj[point] / nl);
  v_cpy(vertex, g);
  return true;
}
}  // namespace

// PQP TriDist, as the reference's in-tree copy C2A/src/C2A.cpp:165-405 minus the
// contact-feature writes; overlap returns 0 (P,Q then hold the last edge pair).
extern "C" double orc_tri_dist(double P[3], double Q[3], const double S[9], const double T[9])
{
  double Sv[9], Tv[9], VEC[3], V[3], Z[3];
  v_sub(&Sv[0], &S[3], &S[0]); v_sub(&Sv[3], &S[6], &S[3]); v_sub(&Sv[6], &S[0], &S[6]);
  v_sub(&Tv[0], &T[3], &T[0]); v_sub(&Tv[3], &T[6], &T[3]); v_sub(&Tv[6], &T[0], &T[6]);

  double minP[3], minQ[3], mindd;
  int shown_disjoint = 0;
  mindd = v_dist2(&S[0], &T[0]) + 1;

  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      orc_seg_points(VEC, P, Q, &S[3 * i], &Sv[3 * i], &T[3 * j], &Tv[3 * j]);
      v_sub(V, Q, P);
      double dd = v_dot(V, V);
      if (dd <= mindd)
      {
        v_cpy(minP, P); v_cpy(minQ, Q); mindd = dd;
        v_sub(Z, &S[3 * ((i + 2) % 3)], P);
        double a = v_dot(Z, VEC);
        v_sub(Z, &T[3 * ((j + 2) % 3)], Q);
        double b = v_dot(Z, VEC);
        if ((a <= 0) && (b >= 0)) return sqrt(dd);
        double p = v_dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
    }

  double onFace[3], vert[3];
  if (face_vertex_case(S, Sv, T, shown_disjoint, onFace, vert))
  {
    v_cpy(P, onFace); v_cpy(Q, vert);
    return sqrt(v_dist2(P, Q));
  }
  if (face_vertex_case(T, Tv, S, shown_disjoint, onFace, vert))
  {
    v_cpy(P, vert); v_cpy(Q, onFace);
    return sqrt(v_dist2(P, Q));
  }
  if (shown_disjoint) { v_cpy(P, minP); v_cpy(Q, minQ); return sqrt(mindd); }
  return 0;
}

// PQP TriDistance (call sites C2A/src/C2A.cpp:1148,1916); same shape as C2A/src/C2A.cpp:408-424.
extern "C" double orc_tri_distance(const double R[9], const double T[3], const double t1[9],
                                   const double t2[9], double p[3], double q[3])
{
  double tri2[9];
  m_v_p(&tri2[0], R, &t2[0], T);
  m_v_p(&tri2[3], R, &t2[3], T);
  m_v_p(&tri2[6], R, &t2[6], T);
  return orc_tri_dist(p, q, t1, tri2);
}

// ---- motion ------------------------------------------------------------------
namespace {
// Matrix3x3::Quaternion_, C2A/LinearMath.h:759-793.  q = (x,y,z,w).
inline void quat_from_matrix(double q[4], const double val[9])
{
  double trace = val[0] + val[4] + val[8];
  if (trace > 0.0)
  {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (val[7] - val[5]) * s;
    q[1] = (val[2] - val[6]) * s;
    q[2] = (val[3] - val[1]) * s;
  }
  else
  {
    int i = val[0] < val[4] ? (val[4] < val[8] ? 2 : 1) : (val[0] < val[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(val[i * 3 + i] - val[j * 3 + j] - val[k * 3 + k] + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (val[k * 3 + j] - val[j * 3 + k]) * s;
    q[j] = (val[j * 3 + i] + val[i * 3 + j]) * s;
    q[k] = (val[k * 3 + i] + val[i * 3 + k]) * s;
  }
}
// Quaternion operator%, C2A/LinearMath.h:1018-1025 (Hamilton product).
inline void quat_mul(double r[4], const double a[4], const double b[4])
{
  r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
// Matrix3x3::Set_Value(Quaternion), C2A/LinearMath.h:809-831.
inline void matrix_from_quat(double v[9], const double q[4])
{
  double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  double s = 2.0 / d;
  double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
  double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
  double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  v[0] = 1.0 - (yy + zz); v[1] = xy - wz; v[2] = xz + wy;
  v[3] = xy + wz; v[4] = 1.0 - (xx + zz); v[5] = yz - wx;
  v[6] = xz - wy; v[7] = yz + wx; v[8] = 1.0 - (xx + yy);
}
}  // namespace

// CInterpMotion ctor (C2A/src/InterpMotion.cpp:148-168) + CInterpMotion_Linear ctor
// (:494-502) -> velocity (:486-491) -> LinearAngularVelocity (:228-270).
extern "C" void orc_motion_init(orc_motion *m, const double R0[9], const double T0[3],
                                const double R1[9], const double T1[3])
{
  memcpy(m->Rs, R0, sizeof(double) * 9); memcpy(m->Ts, T0, sizeof(double) * 3);
  memcpy(m->Re, R1, sizeof(double) * 9); memcpy(m->Te, T1, sizeof(double) * 3);
  memcpy(m->Rc, R0, sizeof(double) * 9); memcpy(m->Tc, T0, sizeof(double) * 3);
  v_sub(m->cv, T1, T0);

  double qs[4], qt[4], q0[4], qd[4];
  quat_from_matrix(qs, m->Rs);
  quat_from_matrix(qt, m->Re);
  q0[0] = -qs[0]; q0[1] = -qs[1]; q0[2] = -qs[2]; q0[3] = qs[3];
  quat_mul(qd, q0, qt);

  double s = 1 < qd[3] ? 1 : qd[3];
  double sign = s < 0 ? -1 : 1;
  double a = (fabs(s - 1) <= 1e-40 || fabs(s + 1) <= 1e-40) ? (2 * sign)
                                                            : (sign * acos(2 * s * s - 1) / sqrt(1 - s * s));
  double tangent[3] = {a * qd[0], a * qd[1], a * qd[2]};
  m->axis[0] = tangent[0]; m->axis[1] = tangent[1]; m->axis[2] = tangent[2];
  m->ang_vel = v_len(tangent);
  double len = m->axis[0] * m->axis[0] + m->axis[1] * m->axis[1] + m->axis[2] * m->axis[2];
  if (len < 0.00000001f) { m->axis[0] = 1.f; m->axis[1] = 0.f; m->axis[2] = 0.f; }
  else
  {
    const double inv = 1.0 / sqrt(len);
    m->axis[0] *= inv; m->axis[1] *= inv; m->axis[2] *= inv;
  }
}

// CInterpMotion_Linear::integrate (C2A/src/InterpMotion.cpp:516-568) with
// AbsoluteRt / DeltaRt (:273-287): mutates the current transform.
extern "C" void orc_motion_integrate(orc_motion *m, double t_in, double q_out[4])
{
  double dt = t_in;
  if (dt > 1) dt = 1;
  m->Tc[0] = m->Ts[0] + dt * m->cv[0];
  m->Tc[1] = m->Ts[1] + dt * m->cv[1];
  m->Tc[2] = m->Ts[2] + dt * m->cv[2];
  double ang = 0.5f * m->ang_vel * dt;
  double sn = sin(ang);
  double drt[4] = {sn * m->axis[0], sn * m->axis[1], sn * m->axis[2], cos(ang)};
  double q0[4], q[4];
  quat_from_matrix(q0, m->Rs);  // re-derived on every call, InterpMotion.cpp:284
  quat_mul(q, q0, drt);
  matrix_from_quat(m->Rc, q);
  if (q_out) { q_out[0] = q[0]; q_out[1] = q[1]; q_out[2] = q[2]; q_out[3] = q[3]; }
}

// CInterpMotion_Linear::computeTOC_MotionBound, C2A/src/InterpMotion.cpp:831-881.
// Normalises N in place, like the reference.
extern "C" double orc_motion_bound_bv(const orc_motion *m, double ang_radius, double N[3])
{
  double cross[3];
  v_normalize(N);
  v_cross(cross, m->axis, N);
  double w_max = (ang_radius)*v_len(cross) * m->ang_vel;
  double v_max = v_dot(m->cv, N);
  if (v_max < 0) v_max = 0;
  double path_max = v_max + w_max;
  if (path_max <= 0) path_max = 1e-30;
  return path_max;
}

// CInterpMotion_Linear::computeTOC(d, r1, S), C2A/src/InterpMotion.cpp:746-828.
extern "C" double orc_motion_bound_leaf(const orc_motion *m, double ang_radius, double S[3])
{
  double v_max, w_max;
  v_normalize(S);
  if (m->ang_vel == 0)
  {
    w_max = 0;
    v_max = v_dot(m->cv, S);
    if (v_max < 0) v_max = 0;
  }
  else
  {
    double cwc[3] = {m->axis[0], m->axis[1], m->axis[2]}, cross[3];
    cwc[0] *= m->ang_vel; cwc[1] *= m->ang_vel; cwc[2] *= m->ang_vel;
    v_cross(cross, cwc, S);
    w_max = ang_radius * v_len(cross);
    v_max = v_dot(m->cv, S);
    if (v_max < 0) v_max = 0;
  }
  double path_max = w_max + v_max;
  if (path_max == 0) path_max = 1e-30;
  return path_max;
}

// ---- traversal + CA loop -----------------------------------------------------
static thread_local std::vector<uint64_t> *g_visits = nullptr;  // non-NULL: toc_recurse logs every visited node pair (orc_solve_visits)
static thread_local orc_spec_stats *g_spec = nullptr;  // non-NULL: exact-mode steps run the speculative split (design study below)
static thread_local double g_spec_prev = -1;
static thread_local orc_replay_stats *g_rp = nullptr;  // non-NULL: every CA step is replayed over the previous step's visit list
static thread_local orc_wide_stats *g_wide = nullptr;
static thread_local double *g_probe = nullptr; static thread_local int g_probe_step = 0;  // design study: CA-loop state when numCA reaches g_probe_step  // non-NULL: exact-mode CA steps run as sequential prefix + level-synchronous wide phase
namespace {
struct Step;
struct RpEntry;
void replay_step(Step &st, const double R[9], const double T[3]);
void spec_step(Step &st, const double R[9], const double T[3]);
void wide_step(Step &st, const double R[9], const double T[3]);
struct Step
{
  const orc_bvh *A, *B;
  const orc_motion *m1, *m2;
  double Rrel[9], Trel[3];  // res->R, res->T
  double abs_err, rel_err, upbound;
  double distance, mint;
  double p1[3], p2[3];
  int num_bv_tests, num_tri_tests;
  int last_a, last_b;
};

inline double bv_size(const orc_bvh *M, int n)
{
  const double *l = &M->l[2 * n];
  return (sqrt(l[0] * l[0] + l[1] * l[1]) + 2 * M->r[n]);  // PQP BV::GetSize, RSS form
}

// C2A_BV_Distance, C2A/src/C2A_BV.cpp:666-675
inline double bv_distance(const double R[9], const double T[3], const orc_bvh *A, int a, const orc_bvh *B,
                          int b, double S[3])
{
  double P[3], Q[3];
  double dist = orc_rect_dist(R, T, &A->l[2 * a], &B->l[2 * b], P, Q, S);
  dist -= (A->r[a] + B->r[b]);
  return (dist < 0.0) ? 0.0 : dist;
}

// TOCStepRecurse_Dis, C2A/src/C2A.cpp:1114-1354
void toc_recurse(Step &st, const double R[9], const double T[3], int b1, int b2)
{
  const orc_bvh *A = st.A, *B = st.B;
  if (g_visits) g_visits->push_back(((uint64_t)(uint32_t)b1 << 32) | (uint32_t)b2);
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  const double *r1 = st.m1->Rc, *tt1 = st.m1->Tc;

  if (l1 && l2)
  {
    double p[3], q[3];
    const double *t1 = &A->tris[9 * (-A->first_child[b1] - 1)];
    const double *t2 = &B->tris[9 * (-B->first_child[b2] - 1)];
    double dTri = orc_tri_distance(st.Rrel, st.Trel, t1, t2, p, q);
    if (dTri <= st.distance)
    {
      if (getenv("ORC_DPROFILE")) fprintf(stderr, "  event at bv %d tri %d: D %.6f -> %.6f\n", st.num_bv_tests, st.num_tri_tests, st.distance, dTri);
      st.distance = dTri;
      double w1[3], w2[3], S1[3], S2[3], tmp[3];
      m_v(tmp, r1, p); v_add(w1, tmp, tt1);
      m_v(tmp, r1, q); v_add(w2, tmp, tt1);
      v_sub(S1, w2, w1);
      S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
      v_cpy(st.p1, p); v_cpy(st.p2, q);
      double mb1 = orc_motion_bound_leaf(st.m1, A->ang_radius[b1], S1);
      double mb2 = orc_motion_bound_leaf(st.m2, B->ang_radius[b2], S2);
      double mint = (dTri) / (mb1 + mb2);
      if (mint < 0.0) mint = 0.0;
      if (mint <= st.mint) st.mint = mint;
      st.last_a = -A->first_child[b1] - 1; st.last_b = -B->first_child[b2] - 1;  // o1->last_tri = t1; o2->last_tri = t2
    }
    st.num_tri_tests++;
    return;
  }

  int a1, a2, c1, c2;
  double R1[9], T1[3], R2[9], T2[3], Tt[3];
  double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    a1 = A->first_child[b1]; a2 = b2; c1 = a1 + 1; c2 = b2;
    mt_m(R1, &A->R[9 * a1], R); v_sub(Tt, T, &A->Tr[3 * a1]); mt_v(T1, &A->R[9 * a1], Tt);
    mt_m(R2, &A->R[9 * c1], R); v_sub(Tt, T, &A->Tr[3 * c1]); mt_v(T2, &A->R[9 * c1], Tt);
  }
  else
  {
    a1 = b1; a2 = B->first_child[b2]; c1 = b1; c2 = a2 + 1;
    m_m(R1, R, &B->R[9 * a2]); m_v_p(T1, R, &B->Tr[3 * a2], T);
    m_m(R2, R, &B->R[9 * c2]); m_v_p(T2, R, &B->Tr[3 * c2], T);
  }

  double S1[3], S2[3], tmp[3], minta, mintb;
  double d1 = bv_distance(R1, T1, A, a1, B, a2, S1);
  if (d1 != 0.0)
  {
    m_v(tmp, &A->R_loc[9 * a1], S1); m_v(S1, r1, tmp);
    S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
    double mb1 = orc_motion_bound_bv(st.m1, A->ang_radius[a1], S1);
    double mb2 = orc_motion_bound_bv(st.m2, B->ang_radius[a2], S2);
    mintb = (d1) / (mb1 + mb2);
    if (mintb <= 0) mintb = 0.0;
  }
  else mintb = 0.0;

  double d2 = bv_distance(R2, T2, A, c1, B, c2, S1);
  if (d2 != 0.0)
  {
    m_v(tmp, &A->R_loc[9 * c1], S1); m_v(S1, r1, tmp);
    S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
    double mb1 = orc_motion_bound_bv(st.m1, A->ang_radius[c1], S1);
    double mb2 = orc_motion_bound_bv(st.m2, B->ang_radius[c2], S2);
    minta = (d2) / (mb1 + mb2);
    if (minta <= 0) minta = 0.0;
  }
  else minta = 0.0;

  st.num_bv_tests += 2;

  // descend test evaluated with the CURRENT st.distance at that moment, :1281-1351
#define DESCEND(mt, d) ((mt) < st.upbound && (((d) < (st.distance - st.abs_err)) || ((d) * (1 + st.rel_err) < st.distance)))
  if (d2 < d1)
  {
    if (DESCEND(minta, d2)) toc_recurse(st, R2, T2, c1, c2);
    else if (minta < st.mint) st.mint = minta;
    if (DESCEND(mintb, d1)) toc_recurse(st, R1, T1, a1, a2);
    else if (mintb < st.mint) st.mint = mintb;
  }
  else
  {
    if (DESCEND(mintb, d1)) toc_recurse(st, R1, T1, a1, a2);
    else if (mintb < st.mint) st.mint = mintb;
    if (DESCEND(minta, d2)) toc_recurse(st, R2, T2, c1, c2);
    else if (minta < st.mint) st.mint = minta;
  }
#undef DESCEND
}

// ---- design study for round 2 (TEST INFRASTRUCTURE): a CA step's traversal split into subtrees that are
// evaluated independently under a GUESSED entry distance and stitched back exactly ------------------------------
// The traversal's control flow depends on the running distance only through comparisons (descend: d < dist;
// leaf: dTri <= dist).  A subtree run records, for as long as its distance is still the entry value, the interval
// of entry values under which every one of those comparisons has the outcome it had; once a leaf inside lowers the
// distance, later comparisons no longer involve the entry value.  If the true entry distance (known when the
// sequential order reaches the subtree) lies in that interval, the recorded outcome -- exit distance, folded step
// bound, counters, last improving triangle pair -- IS the sequential one; otherwise the subtree is re-run.
// Exact mode only (abs_err = rel_err = 0: every step of a long query past its fifth).
struct SpecOut
{
  double exit_dist; bool updated;
  double mint;
  int last_a, last_b; double p1[3], p2[3];
  int nbv, ntri;
  double gt, ge, le, lt;  // valid iff entry > gt && entry >= ge && entry <= le && entry < lt
  long long work;
};

struct SpecRun
{
  Step *st;            // shared read-only context (models, motions, Rrel/Trel, upbound)
  double dist; bool at_entry;
  SpecOut *o;
};

// the two child tests of an expansion: pure functions of (pair, transform, poses), C2A.cpp:1192-1276
struct Kids { int a1, a2, c1, c2; double R1[9], T1[3], R2[9], T2[3], d1, d2, mt1, mt2; };
void expand_pair(const Step &st, const double R[9], const double T[3], int b1, int b2, Kids &k)
{
  const orc_bvh *A = st.A, *B = st.B;
  const double *r1 = st.m1->Rc;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  double Tt[3];
  double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    k.a1 = A->first_child[b1]; k.a2 = b2; k.c1 = k.a1 + 1; k.c2 = b2;
    mt_m(k.R1, &A->R[9 * k.a1], R); v_sub(Tt, T, &A->Tr[3 * k.a1]); mt_v(k.T1, &A->R[9 * k.a1], Tt);
    mt_m(k.R2, &A->R[9 * k.c1], R); v_sub(Tt, T, &A->Tr[3 * k.c1]); mt_v(k.T2, &A->R[9 * k.c1], Tt);
  }
  else
  {
    k.a1 = b1; k.a2 = B->first_child[b2]; k.c1 = b1; k.c2 = k.a2 + 1;
    m_m(k.R1, R, &B->R[9 * k.a2]); m_v_p(k.T1, R, &B->Tr[3 * k.a2], T);
    m_m(k.R2, R, &B->R[9 * k.c2]); m_v_p(k.T2, R, &B->Tr[3 * k.c2], T);
  }
  double S1[3], S2[3], tmp[3];
  k.d1 = bv_distance(k.R1, k.T1, A, k.a1, B, k.a2, S1);
  if (k.d1 != 0.0)
  {
    m_v(tmp, &A->R_loc[9 * k.a1], S1); m_v(S1, r1, tmp);
    S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
    double mb1 = orc_motion_bound_bv(st.m1, A->ang_radius[k.a1], S1);
    double mb2 = orc_motion_bound_bv(st.m2, B->ang_radius[k.a2], S2);
    k.mt1 = (k.d1) / (mb1 + mb2);
    if (k.mt1 <= 0) k.mt1 = 0.0;
  }
  else k.mt1 = 0.0;
  k.d2 = bv_distance(k.R2, k.T2, A, k.c1, B, k.c2, S1);
  if (k.d2 != 0.0)
  {
    m_v(tmp, &A->R_loc[9 * k.c1], S1); m_v(S1, r1, tmp);
    S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
    double mb1 = orc_motion_bound_bv(st.m1, A->ang_radius[k.c1], S1);
    double mb2 = orc_motion_bound_bv(st.m2, B->ang_radius[k.c2], S2);
    k.mt2 = (k.d2) / (mb1 + mb2);
    if (k.mt2 <= 0) k.mt2 = 0.0;
  }
  else k.mt2 = 0.0;
}

void spec_recurse(SpecRun &r, const double R[9], const double T[3], int b1, int b2)
{
  const Step &st = *r.st;
  const orc_bvh *A = st.A, *B = st.B;
  SpecOut &o = *r.o;
  o.work++;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  if (l1 && l2)
  {
    const double *r1 = st.m1->Rc, *tt1 = st.m1->Tc;
    double p[3], q[3];
    const double *t1 = &A->tris[9 * (-A->first_child[b1] - 1)];
    const double *t2 = &B->tris[9 * (-B->first_child[b2] - 1)];
    double dTri = orc_tri_distance(st.Rrel, st.Trel, t1, t2, p, q);
    const bool take = dTri <= r.dist;
    if (r.at_entry)
    {
      if (take) { if (dTri > o.ge) o.ge = dTri; }   // entry >= dTri
      else if (dTri < o.lt) o.lt = dTri;            // entry <  dTri (a NaN dTri is never taken and constrains nothing)
    }
    if (take)
    {
      r.dist = dTri; r.at_entry = false;
      o.exit_dist = dTri; o.updated = true;
      double w1[3], w2[3], S1[3], S2[3], tmp[3];
      m_v(tmp, r1, p); v_add(w1, tmp, tt1);
      m_v(tmp, r1, q); v_add(w2, tmp, tt1);
      v_sub(S1, w2, w1);
      S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
      v_cpy(o.p1, p); v_cpy(o.p2, q);
      double mb1 = orc_motion_bound_leaf(st.m1, A->ang_radius[b1], S1);
      double mb2 = orc_motion_bound_leaf(st.m2, B->ang_radius[b2], S2);
      double mint = (dTri) / (mb1 + mb2);
      if (mint < 0.0) mint = 0.0;
      if (mint <= o.mint) o.mint = mint;
      o.last_a = -A->first_child[b1] - 1; o.last_b = -B->first_child[b2] - 1;
    }
    o.ntri++;
    return;
  }
  Kids k;
  expand_pair(st, R, T, b1, b2, k);
  o.nbv += 2;
  // exact mode: the descend test is mt < upbound && d < dist (d * (1 + 0) and dist - 0 are exact)
  auto child = [&](double d, double mt, const double *Rc, const double *Tc, int n1, int n2) {
    bool go = mt < st.upbound;
    if (go)
    {
      const bool closer = d < r.dist;
      if (r.at_entry)
      {
        if (closer) { if (d > o.gt) o.gt = d; }   // entry >  d
        else if (d < o.le) o.le = d;              // entry <= d
      }
      go = closer;
    }
    if (go) spec_recurse(r, Rc, Tc, n1, n2);
    else if (mt < o.mint) o.mint = mt;
  };
  if (k.d2 < k.d1) { child(k.d2, k.mt2, k.R2, k.T2, k.c1, k.c2); child(k.d1, k.mt1, k.R1, k.T1, k.a1, k.a2); }
  else { child(k.d1, k.mt1, k.R1, k.T1, k.a1, k.a2); child(k.d2, k.mt2, k.R2, k.T2, k.c1, k.c2); }
}

struct SpecTask { double R[9], T[3]; int b1, b2; SpecOut out; bool reached; };
// the top of the tree down to the frontier, evaluated once
struct TopNode { int kid[2]; double d[2], mt[2]; int task; };

int spec_build_top(Step &st, std::vector<TopNode> &top, std::vector<SpecTask> &tasks, const double R[9], const double T[3],
                   int b1, int b2, int depth, int K)
{
  const orc_bvh *A = st.A, *B = st.B;
  TopNode n; n.kid[0] = n.kid[1] = -1; n.task = -1; n.d[0] = n.d[1] = n.mt[0] = n.mt[1] = 0;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  const int id = (int)top.size();
  top.push_back(n);
  if ((l1 && l2) || depth == K)
  {
    SpecTask t; memcpy(t.R, R, 72); memcpy(t.T, T, 24); t.b1 = b1; t.b2 = b2; t.reached = false;
    top[id].task = (int)tasks.size();
    tasks.push_back(t);
    return id;
  }
  Kids k;
  expand_pair(st, R, T, b1, b2, k);
  const bool c_first = k.d2 < k.d1;  // kid[0] = the child visited first
  const double dd[2] = {c_first ? k.d2 : k.d1, c_first ? k.d1 : k.d2}, mm[2] = {c_first ? k.mt2 : k.mt1, c_first ? k.mt1 : k.mt2};
  for (int j = 0; j < 2; j++)
  {
    top[id].d[j] = dd[j]; top[id].mt[j] = mm[j];
    const bool isc = (j == 0) == c_first;
    if (mm[j] < st.upbound)  // a child failing this is never descended, whatever the distance
    {
      const int kid = isc ? spec_build_top(st, top, tasks, k.R2, k.T2, k.c1, k.c2, depth + 1, K)
                          : spec_build_top(st, top, tasks, k.R1, k.T1, k.a1, k.a2, depth + 1, K);
      top[id].kid[j] = kid;
    }
  }
  return id;
}

void spec_stitch(Step &st, std::vector<TopNode> &top, std::vector<SpecTask> &tasks, int id, orc_spec_stats &ss)
{
  if (top[id].task >= 0)
  {
    SpecTask &t = tasks[top[id].task];
    t.reached = true;
    SpecOut &o = t.out;
    const double e = st.distance;
    const bool valid = e > o.gt && e >= o.ge && e <= o.le && e < o.lt;
    ss.reached++;
    if (valid)
    {
      ss.valid++;
      if (o.updated) { st.distance = o.exit_dist; v_cpy(st.p1, o.p1); v_cpy(st.p2, o.p2); st.last_a = o.last_a; st.last_b = o.last_b; }
      if (o.mint < st.mint) st.mint = o.mint;
      st.num_bv_tests += o.nbv; st.num_tri_tests += o.ntri;
      ss.work_seq += o.work;
    }
    else
    {
      const int nbv0 = st.num_bv_tests, ntri0 = st.num_tri_tests;
      toc_recurse(st, t.R, t.T, t.b1, t.b2);
      const long long w = (st.num_bv_tests - nbv0) / 2 + (st.num_tri_tests - ntri0);
      ss.work_seq += w; ss.work_fallback += w;
    }
    return;
  }
  st.num_bv_tests += 2;
  ss.work_top++;
  for (int j = 0; j < 2; j++)
  {
    const bool go = top[id].mt[j] < st.upbound && top[id].d[j] < st.distance;
    if (go) spec_stitch(st, top, tasks, top[id].kid[j], ss);
    else if (top[id].mt[j] < st.mint) st.mint = top[id].mt[j];
  }
}

// one CA step in exact mode; on entry st.distance holds the seed distance and st.mint = 1
void spec_step(Step &st, const double R[9], const double T[3])
{
  orc_spec_stats &ss = *g_spec;
  std::vector<TopNode> top; std::vector<SpecTask> tasks;
  spec_build_top(st, top, tasks, R, T, 0, 0, 0, ss.depth);
  // the guess: the previous exact step's final distance (poses differ by one small advancement)
  const double guess = (g_spec_prev >= 0 && g_spec_prev < st.distance) ? g_spec_prev : st.distance;
  long long spec_total = 0, spec_max = 0;
  for (auto &t : tasks)
  {
    SpecOut &o = t.out;
    o.exit_dist = 0; o.updated = false; o.mint = 1e300; o.last_a = o.last_b = -1; o.nbv = o.ntri = 0; o.work = 0;
    o.p1[0] = o.p1[1] = o.p1[2] = o.p2[0] = o.p2[1] = o.p2[2] = 0;
    o.gt = -1e300; o.ge = -1e300; o.le = 1e300; o.lt = 1e300;
    SpecRun r; r.st = &st; r.dist = guess; r.at_entry = true; r.o = &o;
    spec_recurse(r, t.R, t.T, t.b1, t.b2);
    spec_total += o.work; if (o.work > spec_max) spec_max = o.work;
  }
  const long long fb0 = ss.work_fallback, top0 = ss.work_top;
  spec_stitch(st, top, tasks, 0, ss);
  long long wasted = 0;
  for (auto &t : tasks) if (!t.reached) wasted += t.out.work;
  ss.steps++; ss.tasks += (long long)tasks.size(); ss.work_wasted += wasted;
  if (spec_max > ss.work_max_task) ss.work_max_task = spec_max;
  const long long fb = ss.work_fallback - fb0, tp = ss.work_top - top0;
  const int P[3] = {8, 32, 128};
  for (int j = 0; j < 3; j++)
  {
    const double par = (double)spec_total / P[j];
    ss.par_time[j] += (double)top.size() / 16.0 + (par > (double)spec_max ? par : (double)spec_max) + (double)fb + (double)tp;
  }
  g_spec_prev = st.distance;
}

// ---- design study for round 2 (TEST INFRASTRUCTURE): a CA step replayed over the previous step's visit list -------
// The list keeps, per visited node pair, only its STRUCTURE (ids, parent entry, which child of the parent it is, links
// to its own children entries); the values -- the pair's transform, its two child tests, or the triangle test of a leaf
// pair -- are pure functions of the pair and the current poses and are re-evaluated for the whole list BEFORE the walk
// (on the GPU: level by level across all idle warps; here in list order, parents precede children).  The walk is then
// the reference's depth-first traversal with every evaluation replaced by a record look-up through the parent's child
// link; a pair the previous step did not visit is evaluated on the spot.  Decisions are made with the current running
// distance exactly as in toc_recurse, so the result is the sequential one.
struct RpEntry
{
  int b1, b2, parent, slot, kid[2];
  bool leaf, visited;
  double R[9], T[3];
  Kids k;                              // inner pair: both child tests
  double dTri, p[3], q[3], leaf_mt;    // leaf pair: triangle distance, closest points, the leaf's step bound
};

thread_local std::vector<RpEntry> g_rp_prev;  // the previous step's visit list (structure only is carried over)
void rp_reset() { std::vector<RpEntry>().swap(g_rp_prev); }

void rp_eval(const Step &st, RpEntry &e)
{
  const orc_bvh *A = st.A, *B = st.B;
  if (e.leaf)
  {
    const double *r1 = st.m1->Rc, *tt1 = st.m1->Tc;
    const double *t1 = &A->tris[9 * (-A->first_child[e.b1] - 1)];
    const double *t2 = &B->tris[9 * (-B->first_child[e.b2] - 1)];
    e.dTri = orc_tri_distance(st.Rrel, st.Trel, t1, t2, e.p, e.q);
    double w1[3], w2[3], S1[3], S2[3], tmp[3];
    m_v(tmp, r1, e.p); v_add(w1, tmp, tt1);
    m_v(tmp, r1, e.q); v_add(w2, tmp, tt1);
    v_sub(S1, w2, w1);
    S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
    double mb1 = orc_motion_bound_leaf(st.m1, A->ang_radius[e.b1], S1);
    double mb2 = orc_motion_bound_leaf(st.m2, B->ang_radius[e.b2], S2);
    double mint = (e.dTri) / (mb1 + mb2);
    if (mint < 0.0) mint = 0.0;
    e.leaf_mt = mint;
  }
  else expand_pair(st, e.R, e.T, e.b1, e.b2, e.k);
}

struct RpWalk { Step *st; std::vector<RpEntry> *prev, *cur; orc_replay_stats *ss; int last_prev; };

void rp_visit(RpWalk &w, int pi, const double R[9], const double T[3], int b1, int b2, int parent_new, int slot)
{
  Step &st = *w.st;
  const orc_bvh *A = st.A, *B = st.B;
  RpEntry ne;
  ne.b1 = b1; ne.b2 = b2; ne.parent = parent_new; ne.slot = slot; ne.kid[0] = ne.kid[1] = -1; ne.visited = false;
  ne.leaf = A->first_child[b1] < 0 && B->first_child[b2] < 0;
  const int ni = (int)w.cur->size();
  if (parent_new >= 0) (*w.cur)[parent_new].kid[slot] = ni;
  w.ss->visits++;
  if (pi >= 0)
  {
    RpEntry &pe = (*w.prev)[pi];
    pe.visited = true;
    w.ss->hits++;
    if (pi == w.last_prev + 1) w.ss->seq_hits++;
    w.last_prev = pi;
    // take the pre-evaluated values
    ne.k = pe.k; ne.dTri = pe.dTri; ne.leaf_mt = pe.leaf_mt;
    v_cpy(ne.p, pe.p); v_cpy(ne.q, pe.q);
  }
  else
  {
    w.ss->misses++;
    memcpy(ne.R, R, 72); memcpy(ne.T, T, 24);
    rp_eval(st, ne);
  }
  w.cur->push_back(ne);
  if (ne.leaf)
  {
    // C2A.cpp:1141-1183 with the pure parts already evaluated
    if (ne.dTri <= st.distance)
    {
      st.distance = ne.dTri;
      v_cpy(st.p1, ne.p); v_cpy(st.p2, ne.q);
      if (ne.leaf_mt <= st.mint) st.mint = ne.leaf_mt;
      st.last_a = -A->first_child[b1] - 1; st.last_b = -B->first_child[b2] - 1;
    }
    st.num_tri_tests++;
    return;
  }
  const Kids k = ne.k;  // (a copy: w.cur may grow while we recurse)
  st.num_bv_tests += 2;
  const bool c_first = k.d2 < k.d1;
  for (int j = 0; j < 2; j++)
  {
    const bool isc = (j == 0) == c_first;
    const double d = isc ? k.d2 : k.d1, mt = isc ? k.mt2 : k.mt1;
    const int s = isc ? 1 : 0;
    const bool go = mt < st.upbound && ((d < (st.distance - st.abs_err)) || (d * (1 + st.rel_err) < st.distance));
    if (go)
      rp_visit(w, pi >= 0 ? (*w.prev)[pi].kid[s] : -1, isc ? k.R2 : k.R1, isc ? k.T2 : k.T1, isc ? k.c1 : k.a1, isc ? k.c2 : k.a2, ni, s);
    else if (mt < st.mint) st.mint = mt;
  }
}

void replay_step(Step &st, const double R[9], const double T[3])
{
  orc_replay_stats &ss = *g_rp;
  std::vector<RpEntry> prev;
  prev.swap(g_rp_prev);
  // pre-evaluation of the whole previous list for the current poses (parents precede children in visit order)
  for (size_t i = 0; i < prev.size(); i++)
  {
    RpEntry &e = prev[i];
    e.visited = false;
    if (e.parent < 0) { memcpy(e.R, R, 72); memcpy(e.T, T, 24); }
    else
    {
      const Kids &pk = prev[e.parent].k;
      memcpy(e.R, e.slot ? pk.R2 : pk.R1, 72); memcpy(e.T, e.slot ? pk.T2 : pk.T1, 24);
    }
    rp_eval(st, e);
    ss.preeval++;
  }
  std::vector<RpEntry> cur;
  cur.reserve(prev.size() + 64);
  RpWalk w; w.st = &st; w.prev = &prev; w.cur = &cur; w.ss = &ss; w.last_prev = -1;
  rp_visit(w, prev.empty() ? -1 : 0, R, T, 0, 0, -1, 0);
  for (auto &e : prev) if (!e.visited) ss.wasted++;
  ss.steps++;
  g_rp_prev.swap(cur);
}

// ---- round-2 algorithm (TEST INFRASTRUCTURE: the CPU statement of what c2a_wide.cuh does on the device) ---------
// One exact-mode CA step (abs_err = rel_err = 0, so the descend test is "mt < upbound && d < dist") as a depth-first
// traversal that pops W node pairs at a time:
//   * the pending pairs live on a stack ordered by the reference's visiting order (top = visited next).  A round pops
//     the top W pairs, runs their child tests (or the triangle test of a leaf pair) side by side, and pushes the
//     children back in visiting order (of one pair: the closer child above the other; of the window: the first
//     pair's children on top) -- so the stack stays in visiting order without any sorting;
//   * every node carries a pre-order key (one bit per level: 0 = the child visited first) and M = the maximum, over
//     its path from the root, of its ancestors' and its own test value d (+inf where mt >= upbound);
//   * the running distance changes only at leaves with dTri <= dist ("events").  Everything that precedes the top
//     of the stack has been evaluated, so the events up to there can be resolved exactly, in key order: Dw, the
//     distance in force at the top of the stack.  A popped pair is expanded iff its value < Dw and its children are
//     pushed iff theirs are: the distance only shrinks, so this is a superset of what the reference expands; the
//     pairs of one window are expanded without regard to the events inside the window (the speculation);
//   * when the stack is empty all events are known.  FOLD: with D(key) the distance in force at a node's pre-order
//     position, a node was visited iff M < D(key); visited parents count their two child tests, children that fail
//     their own test at their own position fold their step bound, visited leaves count a triangle test, event
//     leaves fold theirs and the last one gives distance, p1/p2 and last_tri.
// "visited <=> M(n) < D(key(n))": => is immediate (D never grows).  <= could only fail if some ancestor test value
// of an event leaf l is >= dTri(l): that edge passed under an earlier, larger distance and descendants visited
// after l would be mis-classified.  A bounding-volume distance never exceeds the distance of the triangles inside
// up to rounding, so this is rare; it is DETECTED (M of the leaf's parent >= dTri, dTri != 0) and the step is then
// redone sequentially.  With that check the result is the sequential one, bit for bit; W = 1 is the sequential walk.
struct WNode
{
  uint64_t key[2]; int depth;    // path bits, left-aligned (bit 127 = first level)
  double Mpar, val, mt;          // M of the parent (-inf for the root), own test value (d, or +inf if mt >= upbound), own step bound
  int b1, b2; bool leafpair, expanded;
  double R[9], T[3];
  double dTri, p[3], q[3], leaf_mt;  // leaf pairs that were evaluated
};
constexpr int WKEY_BITS = 128;
inline void wkey_child(uint64_t out[2], const uint64_t key[2], int depth, int bit)  // depth = the parent's
{
  out[0] = key[0]; out[1] = key[1];
  if (bit) out[depth >> 6] |= (uint64_t)1 << (63 - (depth & 63));
}
inline void wkey_parent(uint64_t out[2], const uint64_t key[2], int depth)  // depth = the child's (>= 1)
{
  out[0] = key[0]; out[1] = key[1];
  out[(depth - 1) >> 6] &= ~((uint64_t)1 << (63 - ((depth - 1) & 63)));
}
inline bool wkey_less(const uint64_t a[2], const uint64_t b[2]) { return a[0] < b[0] || (a[0] == b[0] && a[1] < b[1]); }

inline void wide_leaf_eval(const Step &st, WNode &n)
{
  const orc_bvh *A = st.A, *B = st.B;
  const double *r1 = st.m1->Rc, *tt1 = st.m1->Tc;
  const double *t1 = &A->tris[9 * (-A->first_child[n.b1] - 1)];
  const double *t2 = &B->tris[9 * (-B->first_child[n.b2] - 1)];
  n.dTri = orc_tri_distance(st.Rrel, st.Trel, t1, t2, n.p, n.q);
  double w1[3], w2[3], S1[3], S2[3], tmp[3];
  m_v(tmp, r1, n.p); v_add(w1, tmp, tt1);
  m_v(tmp, r1, n.q); v_add(w2, tmp, tt1);
  v_sub(S1, w2, w1);
  S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
  double mb1 = orc_motion_bound_leaf(st.m1, A->ang_radius[n.b1], S1);
  double mb2 = orc_motion_bound_leaf(st.m2, B->ang_radius[n.b2], S2);
  double mint = (n.dTri) / (mb1 + mb2);
  if (mint < 0.0) mint = 0.0;
  n.leaf_mt = mint;
}

void wide_step(Step &st, const double R[9], const double T[3])
{
  orc_wide_stats &ws = *g_wide;
  const orc_bvh *A = st.A, *B = st.B;
  const Step st0 = st;  // for the sequential redo
  const double INF = std::numeric_limits<double>::infinity();
  const int W = (int)ws.window;
  ws.steps++;

  std::vector<WNode> nodes;     // every evaluated node, in evaluation order (the records of the fold)
  std::vector<int> stack;       // indices into nodes; back() = visited next
  std::vector<int> unresolved;  // evaluated leaves whose event status is still open
  std::vector<int> eleaf;       // event leaves, in key order
  double Dw = st.distance;      // the distance in force at the top of the stack
  bool anomaly = false, overflow = false;
  {
    WNode n; n.key[0] = n.key[1] = 0; n.depth = 0; n.Mpar = -INF; n.val = -INF; n.mt = 0; n.expanded = false;
    n.b1 = 0; n.b2 = 0; memcpy(n.R, R, 72); memcpy(n.T, T, 24);
    n.leafpair = A->first_child[0] < 0 && B->first_child[0] < 0;
    nodes.push_back(n); stack.push_back(0);
  }
  long long rounds = 0, leaf_passes = 0;
  std::vector<int> win, kids;
  std::vector<int> pending;     // leaf pairs waiting for a LEAF pass (device: batched, 32 lanes per pass); empty when leaf_batch = 0
  const int LB = (int)ws.leaf_batch;
  while ((!stack.empty() || !pending.empty()) && !anomaly && !overflow)
  {
    kids.clear();
    if (LB > 0 && ((int)pending.size() >= LB || stack.empty()))
    {
      // LEAF pass: up to 32 waiting leaf pairs, oldest first; those that no longer pass are dropped
      leaf_passes++;
      int done = 0; size_t i = 0;
      for (; i < pending.size() && done < 32; i++)
      {
        WNode &n = nodes[pending[i]];
        const double M = n.Mpar > n.val ? n.Mpar : n.val;
        if (!(M < Dw)) continue;
        n.expanded = true; wide_leaf_eval(st, n); unresolved.push_back(pending[i]); ws.wide_leaves++; done++;
      }
      pending.erase(pending.begin(), pending.begin() + i);
    }
    else
    {
    rounds++;
    // pop the window: entries that fail under Dw are dropped (their step bound is folded at the end, from the records)
    win.clear();
    while (!stack.empty() && (int)win.size() < W)
    {
      const int i = stack.back(); stack.pop_back();
      const double M = nodes[i].Mpar > nodes[i].val ? nodes[i].Mpar : nodes[i].val;
      if (M < Dw) win.push_back(i);
    }
    if ((long long)win.size() > ws.max_width) ws.max_width = (long long)win.size();
    for (int wi : win)
    {
      nodes[wi].expanded = true;
      if (nodes[wi].leafpair) { wide_leaf_eval(st, nodes[wi]); unresolved.push_back(wi); ws.wide_leaves++; continue; }
      if (nodes[wi].depth + 1 > WKEY_BITS) { overflow = true; break; }
      Kids k;
      expand_pair(st, nodes[wi].R, nodes[wi].T, nodes[wi].b1, nodes[wi].b2, k);
      ws.wide_tests += 2;
      const bool c_first = k.d2 < k.d1;
      const double Mn = nodes[wi].Mpar > nodes[wi].val ? nodes[wi].Mpar : nodes[wi].val;
      int kid_idx[2];
      for (int j = 0; j < 2; j++)  // j = 0: the child visited first
      {
        const bool isc = (j == 0) == c_first;
        WNode c; c.depth = nodes[wi].depth + 1; c.expanded = false;
        wkey_child(c.key, nodes[wi].key, nodes[wi].depth, j);
        c.Mpar = Mn;
        const double d = isc ? k.d2 : k.d1, mt = isc ? k.mt2 : k.mt1;
        c.val = (mt < st.upbound) ? d : INF; c.mt = mt;
        c.b1 = isc ? k.c1 : k.a1; c.b2 = isc ? k.c2 : k.a2;
        memcpy(c.R, isc ? k.R2 : k.R1, 72); memcpy(c.T, isc ? k.T2 : k.T1, 24);
        c.leafpair = A->first_child[c.b1] < 0 && B->first_child[c.b2] < 0;
        nodes.push_back(c);
        kid_idx[j] = (int)nodes.size() - 1;
      }
      kids.push_back(kid_idx[0]); kids.push_back(kid_idx[1]);
    }
    // push the children in reverse visiting order (the window's first pair's first child ends up on top)
    for (int i = (int)kids.size() - 1; i >= 0; i--)
      if (nodes[kids[i]].val < Dw)
      {
        if (LB > 0 && nodes[kids[i]].leafpair) pending.push_back(kids[i]); else stack.push_back(kids[i]);
      }
    if ((long long)stack.size() > ws.max_stack) ws.max_stack = (long long)stack.size();
    }
    // resolve the events that precede everything still to be evaluated (the top of the stack, the waiting leaf pairs), in key order
    const uint64_t *front = stack.empty() ? nullptr : nodes[stack.back()].key;
    for (int pi : pending) if (!front || wkey_less(nodes[pi].key, front)) front = nodes[pi].key;
    std::vector<int> ready, later;
    for (int li : unresolved)
      if (!front || wkey_less(nodes[li].key, front)) ready.push_back(li); else later.push_back(li);
    std::sort(ready.begin(), ready.end(), [&](int x, int y) { return wkey_less(nodes[x].key, nodes[y].key); });
    if ((long long)later.size() > ws.max_unresolved) ws.max_unresolved = (long long)later.size();
    for (int li : ready)
    {
      const WNode &n = nodes[li];
      const double M = n.Mpar > n.val ? n.Mpar : n.val;
      if (M < Dw && n.dTri <= Dw)
      {
        if (n.dTri != 0.0 && !(n.Mpar < n.dTri)) { anomaly = true; break; }
        Dw = n.dTri; eleaf.push_back(li);
      }
    }
    unresolved.swap(later);
  }
  ws.leaf_passes += leaf_passes;
  ws.rounds += rounds;
  if (anomaly || overflow) { ws.anomalies += anomaly ? 1 : 0; ws.redo++; st = st0; toc_recurse(st, R, T, 0, 0); return; }
  ws.events += (long long)eleaf.size();

  // ---- fold
  auto D_at = [&](const uint64_t key[2]) {
    double d = st0.distance;
    for (size_t i = 0; i < eleaf.size(); i++) if (wkey_less(nodes[eleaf[i]].key, key)) d = nodes[eleaf[i]].dTri; else break;
    return d;
  };
  for (const WNode &n : nodes)
  {
    bool reached = true;
    if (n.depth > 0)
    {
      uint64_t pk[2]; wkey_parent(pk, n.key, n.depth);
      reached = n.Mpar < D_at(pk);
      if (reached) { st.num_bv_tests += 1; ws.wide_tests_visited += 1; }
    }
    if (!reached) continue;
    const bool pass = n.val < D_at(n.key);
    if (!pass) { if (n.mt < st.mint) st.mint = n.mt; continue; }
    if (!n.expanded) { ws.redo++; ws.closure_fail++; st = st0; toc_recurse(st, R, T, 0, 0); return; }  // cannot happen (superset); checked all the same
    if (n.leafpair) { st.num_tri_tests++; ws.wide_leaves_visited++; }
  }
  for (size_t i = 0; i < eleaf.size(); i++)
  {
    const WNode &n = nodes[eleaf[i]];
    if (n.leaf_mt <= st.mint) st.mint = n.leaf_mt;
  }
  if (!eleaf.empty())
  {
    const WNode &n = nodes[eleaf.back()];
    st.distance = n.dTri; v_cpy(st.p1, n.p); v_cpy(st.p2, n.q);
    st.last_a = -A->first_child[n.b1] - 1; st.last_b = -B->first_child[n.b2] - 1;
  }
}

// C2A_TimeOfContactStep (rotational branch), C2A/src/C2A.cpp:1778-1931.
// numCA / prev_mint carry res->numCA and the previous step's res->mint.
void toc_step(Step &st, int numCA, int seedA, int seedB)
{
  const orc_bvh *A = st.A, *B = st.B;
  const double *R1 = st.m1->Rc, *T1 = st.m1->Tc, *R2 = st.m2->Rc, *T2 = st.m2->Tc;
  double Tt[3], Rt[9], R[9], T[3];
  mt_m(st.Rrel, R1, R2);
  v_sub(Tt, T2, T1);
  mt_v(st.Trel, R1, Tt);
  m_m(Rt, st.Rrel, &B->R[0]);
  mt_m(R, &A->R[0], Rt);
  m_v_p(Tt, st.Rrel, &B->Tr[0], st.Trel);
  v_sub(Tt, Tt, &A->Tr[0]);
  mt_v(T, &A->R[0], Tt);

  double p[3], q[3];
  st.distance = orc_tri_distance(st.Rrel, st.Trel, &A->tris[9 * seedA], &B->tris[9 * seedB], p, q);
  if (numCA == 0) st.mint = 1;
  if (st.mint <= 0.005 || st.distance <= 0.5 || numCA > 5) { st.abs_err = 0; st.rel_err = 0; }
  else
  {
    st.abs_err = 1e+30;
    st.rel_err = (numCA <= 2) ? 3 : 0.5;
  }
  st.mint = 1;
  if (getenv("ORC_DPROFILE")) fprintf(stderr, "step numCA %d seed dist %.6f abs_err %g (bv so far %d)\n", numCA, st.distance, st.abs_err, st.num_bv_tests);
  if (g_visits) g_visits->push_back(~0ull);  // step separator
  if (g_rp) replay_step(st, R, T);
  else if (g_wide && st.abs_err == 0 && st.rel_err == 0) wide_step(st, R, T);
  else if (g_spec && st.abs_err == 0 && st.rel_err == 0) spec_step(st, R, T);
  else toc_recurse(st, R, T, 0, 0);
}
// ---- translation-only branch (both angular speeds < 1e-8, C2A.cpp:2391-2395) -------------------
// Only objmotion1's velocity enters the inner advancement (the reference calls objmotion1->CAonRSS /
// CAonNonAdjacentTriangles and never reads objmotion2's cv there): reproduced, not fixed.
//
// Undefined behaviour in the reference, and what this restatement does about it: C2ARectDist leaves S
// untouched when no edge pair is accepted and both face separations are negative (the rectangles cross,
// C2A_RectDist.h:887-928), and CAonRSS then reads its uninitialised local S (InterpMotion.cpp:579-592).
// With any finite non-zero garbage the outcome is the same -- d = 0, tocf = -0.1*delta/path_max < 0, the
// call returns true with a negative step, and the caller descends iff 0 < res->distance -- so S is
// preset to (1,0,0) here.  (Garbage that normalises to NaN would make the reference skip that subtree.)
constexpr double SECURITY_RATIO = 0.1;  // GMP_CCD_SECURITY_DISTANCE_RATIO, InterpMotion.cpp:289

// CInterpMotion_Linear::CAonRSS, C2A/src/InterpMotion.cpp:570-655 (bValid_C_clo is always true,
// C2A_RectDist.h:931, so the centre-of-mass branch :607-621 is dead).
bool ca_on_rss(const orc_motion *m1, double delta, const double r1[9], const double R[9], const double T[3],
               const orc_bvh *A, int a, const orc_bvh *B, int b, double *mint, double *distance)
{
  double S[3] = {1, 0, 0}, temp1[3], temp2[3], Vel[3];
  double d = bv_distance(R, T, A, a, B, b, S);
  m_v(temp1, &A->R_loc[9 * a], S);
  m_v(S, r1, temp1);
  const double *cvc = m1->cv;
  mt_v(temp1, r1, cvc);
  mt_v(Vel, &A->R_loc[9 * a], temp1);
  double tocf = (d - SECURITY_RATIO * delta) / orc_motion_bound_leaf(m1, A->ang_radius[a], S);
  double total_toc = tocf;
  int nIters = 1;
  while ((d >= delta) && (total_toc <= mint[0]) && (nIters < 50))
  {
    v_madd(temp2, T, Vel, -total_toc);
    d = bv_distance(R, temp2, A, a, B, b, S);
    if (d == 0) break;
    m_v(temp1, &A->R_loc[9 * a], S);
    m_v(S, r1, temp1);
    tocf = (d - SECURITY_RATIO * delta) / orc_motion_bound_leaf(m1, A->ang_radius[a], S);
    nIters++;
    if (tocf < delta) break;
    total_toc += tocf;
  }
  if (total_toc < 1.0)
  {
    mint[0] = total_toc;
    distance[0] = (total_toc - delta) * v_len(cvc);
    if (distance[0] < 0) distance[0] = 0;
    return true;
  }
  return false;
}

// CInterpMotion_Linear::CAonNonAdjacentTriangles, C2A/src/InterpMotion.cpp:657-743
bool ca_on_triangles(const orc_motion *m1, double delta, const double r1[9], const double triA[9],
                     const double triB[9], double *mint, double *distance)
{
  double p[3], q[3], triA_t[9], Vel[3], n[3];
  const double *cvc = m1->cv;
  mt_v(Vel, r1, cvc);
  double d = orc_tri_dist(p, q, triA, triB);
  v_sub(n, q, p);
  v_normalize(n);
  double u = v_dot(Vel, n);
  if (u <= 0) u = 1e-30;
  if (d == 0) { mint[0] = 0.0; distance[0] = 0.0; return true; }
  double dt = (d - SECURITY_RATIO * delta) / u;
  int nIters = 1;
  if (dt >= mint[0]) return false;
  double total_toc = dt;
  while ((d > delta) && (total_toc <= mint[0]) && (nIters < 50))
  {
    for (int i = 0; i < 3; i++) v_madd(&triA_t[3 * i], &triA[3 * i], Vel, total_toc);
    d = orc_tri_dist(p, q, triA_t, triB);
    if (d == 0.0) break;
    v_sub(n, q, p);
    v_normalize(n);
    u = v_dot(Vel, n);
    if (u <= 0) u = 1e-30;
    double tofc = (d - SECURITY_RATIO * delta) / u;
    nIters++;
    if (tofc < delta) break;
    total_toc += tofc;
  }
  if (total_toc <= mint[0] && total_toc >= 0.0)
  {
    mint[0] = total_toc;
    distance[0] = (total_toc)*v_len(cvc) + delta;
    return true;
  }
  return false;
}

struct TStep
{
  const orc_bvh *A, *B;
  const orc_motion *m1;
  double delta;
  double Rrel[9], Trel[3];
  double distance, mint;
  int num_bv_tests, num_tri_tests;
  int last_a, last_b;  // res->last_triA / last_triB
};

// TOCStepRecurse_Dis_Translation, C2A/src/C2A.cpp:1362-1521 (abs_err = rel_err = 0, :1902-1903)
void toc_recurse_translation(TStep &st, const double R[9], const double T[3], int b1, int b2)
{
  const orc_bvh *A = st.A, *B = st.B;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  const double *r1 = st.m1->Rc;
  if (l1 && l2)
  {
    const int ta = -A->first_child[b1] - 1, tb = -B->first_child[b2] - 1;
    const double *t1 = &A->tris[9 * ta], *t2 = &B->tris[9 * tb];
    double tri2[9];
    m_v_p(&tri2[0], st.Rrel, &t2[0], st.Trel); m_v_p(&tri2[3], st.Rrel, &t2[3], st.Trel); m_v_p(&tri2[6], st.Rrel, &t2[6], st.Trel);
    if (ca_on_triangles(st.m1, st.delta, r1, t1, tri2, &st.mint, &st.distance)) { st.last_a = ta; st.last_b = tb; }
    st.num_tri_tests++;
    return;
  }
  int a1, a2, c1, c2;
  double R1[9], T1[3], R2[9], T2[3], Tt[3];
  double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    a1 = A->first_child[b1]; a2 = b2; c1 = a1 + 1; c2 = b2;
    mt_m(R1, &A->R[9 * a1], R); v_sub(Tt, T, &A->Tr[3 * a1]); mt_v(T1, &A->R[9 * a1], Tt);
    mt_m(R2, &A->R[9 * c1], R); v_sub(Tt, T, &A->Tr[3 * c1]); mt_v(T2, &A->R[9 * c1], Tt);
  }
  else
  {
    a1 = b1; a2 = B->first_child[b2]; c1 = b1; c2 = a2 + 1;
    m_m(R1, R, &B->R[9 * a2]); m_v_p(T1, R, &B->Tr[3 * a2], T);
    m_m(R2, R, &B->R[9 * c2]); m_v_p(T2, R, &B->Tr[3 * c2], T);
  }
  double d1 = 1e+30, d2 = 1e+30, minta = st.mint, mintc = st.mint;
  ca_on_rss(st.m1, st.delta, r1, R1, T1, A, a1, B, a2, &minta, &d1);
  ca_on_rss(st.m1, st.delta, r1, R2, T2, A, c1, B, c2, &mintc, &d2);
  st.num_bv_tests += 2;
#define DESCEND_T(mt, d) ((mt) < st.mint && (((d) < (st.distance - 0.0)) || ((d) * (1 + 0.0) < st.distance)))
  if (d2 < d1)
  {
    if (DESCEND_T(mintc, d2)) toc_recurse_translation(st, R2, T2, c1, c2);
    if (DESCEND_T(minta, d1)) toc_recurse_translation(st, R1, T1, a1, a2);
  }
  else
  {
    if (DESCEND_T(minta, d1)) toc_recurse_translation(st, R1, T1, a1, a2);
    if (DESCEND_T(mintc, d2)) toc_recurse_translation(st, R2, T2, c1, c2);
  }
#undef DESCEND_T
}

// The translation-only path of C2A_TimeOfContactStep (C2A.cpp:1818-1852, :1900-1917) and of
// C2A_QueryTimeOfContact (:2032-2049), with the pose outputs of C2A_Solve (:2411-2429).
void solve_translation(const orc_bvh *A, const orc_bvh *B, orc_motion &m1, orc_motion &m2, int seedA, int seedB,
                       double delta, orc_result *out)
{
  TStep st;
  st.A = A; st.B = B; st.m1 = &m1; st.delta = delta;
  st.num_bv_tests = 0; st.num_tri_tests = 0; st.last_a = seedA; st.last_b = seedB;
  double Tt[3], Rt[9], R[9], T[3];
  mt_m(st.Rrel, m1.Rc, m2.Rc);
  v_sub(Tt, m2.Tc, m1.Tc);
  mt_v(st.Trel, m1.Rc, Tt);
  m_m(Rt, st.Rrel, &B->R[0]);
  mt_m(R, &A->R[0], Rt);
  m_v_p(Tt, st.Rrel, &B->Tr[0], st.Trel);
  v_sub(Tt, Tt, &A->Tr[0]);
  mt_v(T, &A->R[0], Tt);

  const double *t1 = &A->tris[9 * seedA], *t2 = &B->tris[9 * seedB];
  double tri2[9], mint = 1.0, dTri = 0;
  m_v_p(&tri2[0], st.Rrel, &t2[0], st.Trel); m_v_p(&tri2[3], st.Rrel, &t2[3], st.Trel); m_v_p(&tri2[6], st.Rrel, &t2[6], st.Trel);
  if (ca_on_triangles(&m1, delta, m1.Rc, t1, tri2, &mint, &dTri)) { st.mint = mint; st.distance = dTri; }
  else { st.mint = 1.0; st.distance = 1e+30; }
  toc_recurse_translation(st, R, T, 0, 0);

  // :1907-1916: distance of the last improving triangle pair at the pose of the step bound
  orc_motion_integrate(&m1, st.mint, 0);
  orc_motion_integrate(&m2, st.mint, 0);
  double Rrel[9], Trel[3], p[3], q[3];
  mt_m(Rrel, m1.Rc, m2.Rc);
  v_sub(Tt, m2.Tc, m1.Tc);
  mt_v(Trel, m1.Rc, Tt);
  st.distance = orc_tri_distance(Rrel, Trel, &A->tris[9 * st.last_a], &B->tris[9 * st.last_b], p, q);

  out->toc = st.mint;
  out->collisionfree = (st.mint >= 1.0) ? 1 : 0;
  out->numCA = 0;
  out->num_bv_tests = st.num_bv_tests;
  out->num_tri_tests = st.num_tri_tests;
  out->distance = st.distance;
  out->mint = st.mint;
  out->last_tri_a = st.last_a; out->last_tri_b = st.last_b;
  // res->p1 / p2 are copied from never-written locals in the reference (C2A.cpp:1392,1411-1412): left zero
  if (!out->collisionfree)
  {
    orc_motion_integrate(&m1, out->toc, 0);
    orc_motion_integrate(&m2, out->toc, 0);
    memcpy(&out->pose_toc[0], m1.Rc, sizeof(double) * 9); memcpy(&out->pose_toc[9], m1.Tc, sizeof(double) * 3);
    memcpy(&out->pose_toc[12], m2.Rc, sizeof(double) * 9); memcpy(&out->pose_toc[21], m2.Tc, sizeof(double) * 3);
  }
}
}  // namespace

// C2A_Solve (C2A/src/C2A.cpp:2315-2444) up to and including the pose outputs,
// around C2A_QueryTimeOfContact (:1987-2146); the contact pass (:2433) is not
// part of the parity contract.
extern "C" void orc_solve(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA,
                          int32_t seedB, double tol_d, double tol_t, orc_result *out)
{
  orc_motion m1, m2;
  orc_motion_init(&m1, &poses[0], &poses[9], &poses[12], &poses[21]);
  orc_motion_init(&m2, &poses[24], &poses[33], &poses[36], &poses[45]);
  memset(out, 0, sizeof(*out));
  if (m1.ang_vel < 1e-8 && m2.ang_vel < 1e-8)
  {
    // translation-only branch (C2A.cpp:2391-2395, :1362-1521); m_toc_delta = d_delta (:2387-2388)
    solve_translation(A, B, m1, m2, seedA, seedB, tol_d, out);
    return;
  }

  Step st;
  st.A = A; st.B = B; st.m1 = &m1; st.m2 = &m2;
  st.num_bv_tests = 0; st.num_tri_tests = 0; st.upbound = 1; st.mint = 1; st.distance = 0;
  st.last_a = -1; st.last_b = -1;
  st.p1[0] = st.p1[1] = st.p1[2] = st.p2[0] = st.p2[1] = st.p2[2] = 0;
  int numCA = 0;
  double lamda = 0.0, lastLamda = 0, dlamda = 0.0, toc = 0;
  int collisionfree = 0;
  bool done_free = false;

  toc_step(st, numCA, seedA, seedB);
  int nItrs = 0;
  double dist = st.distance, mint = st.mint;
  numCA = 1;
  lastLamda = mint;

  while (dist > tol_d)
  {
    nItrs++;
    if (nItrs > 150) break;
    if (mint >= 1.0) { collisionfree = 1; toc = 0; done_free = true; break; }
    dlamda = mint;
    if (dlamda < tol_t) break;
    lamda += dlamda;
    if (lamda >= 1.0) { collisionfree = 1; toc = 0; done_free = true; break; }
    lastLamda = lamda;
    numCA++;
    if (g_probe && numCA == g_probe_step) { g_probe[0] = lamda; g_probe[1] = dist; g_probe[2] = mint; g_probe[3] = st.num_bv_tests; g_probe[4] = st.num_tri_tests; }
    orc_motion_integrate(&m1, lamda, 0);
    orc_motion_integrate(&m2, lamda, 0);
    st.upbound = 1.0 - lamda;
    toc_step(st, numCA, seedA, seedB);
    dist = st.distance;
    mint = st.mint;
  }

  if (!done_free)
  {
    if (dist == 0 && mint < 0) toc = lastLamda - dlamda;
    else toc = lastLamda;
    collisionfree = 0;
    if (toc >= 1 - tol_t) toc = 0;
    // pose at toc: C2A.cpp:2143 and :2411-2429 (trans0/trans1 = integrate(toc))
    orc_motion_integrate(&m1, toc, 0);
    orc_motion_integrate(&m2, toc, 0);
    memcpy(&out->pose_toc[0], m1.Rc, sizeof(double) * 9); memcpy(&out->pose_toc[9], m1.Tc, sizeof(double) * 3);
    memcpy(&out->pose_toc[12], m2.Rc, sizeof(double) * 9); memcpy(&out->pose_toc[21], m2.Tc, sizeof(double) * 3);
  }
  out->collisionfree = collisionfree;
  out->numCA = numCA;
  out->num_bv_tests = st.num_bv_tests;
  out->num_tri_tests = st.num_tri_tests;
  out->toc = toc;
  out->distance = st.distance;
  out->mint = st.mint;
  v_cpy(out->p1, st.p1); v_cpy(out->p2, st.p2);
  out->last_tri_a = st.last_a; out->last_tri_b = st.last_b;
}

extern "C" void orc_solve_batch(const orc_bvh *A, const orc_bvh *B, const double *poses, int64_t n,
                                const int32_t *seedA, const int32_t *seedB, double tol_d, double tol_t,
                                orc_result *out, int32_t n_threads)
{
  if (n_threads < 1) n_threads = 1;
  // dynamic queue of 16-query chunks: the cost per query has a long tail, a static split would time the unluckiest thread
  std::atomic<int64_t> next(0);
  auto work = [&](int) {
    while (true)
    {
      const int64_t lo = next.fetch_add(16);
      if (lo >= n) break;
      const int64_t hi = lo + 16 < n ? lo + 16 : n;
      for (int64_t i = lo; i < hi; i++)
        orc_solve(A, B, poses + 48 * i, seedA ? seedA[i] : 0, seedB ? seedB[i] : 0, tol_d, tol_t, &out[i]);
    }
  };
  if (n_threads == 1) { work(0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
  for (auto &t : th) t.join();
}

// ---- contact pass ------------------------------------------------------------------------
// TriDist with contact features: the reference's in-tree copy, C2A/src/C2A.cpp:165-405.
extern "C" double orc_tri_dist_features(double P[3], double Q[3], const double S[9], const double T[9],
                                        int32_t *f1_type, int32_t *f1_fid, int32_t *f2_type, int32_t *f2_fid,
                                        int32_t *collided)
{
  double Sv[9], Tv[9], VEC[3], V[3], Z[3];
  v_sub(&Sv[0], &S[3], &S[0]); v_sub(&Sv[3], &S[6], &S[3]); v_sub(&Sv[6], &S[0], &S[6]);
  v_sub(&Tv[0], &T[3], &T[0]); v_sub(&Tv[3], &T[6], &T[3]); v_sub(&Tv[6], &T[0], &T[6]);
  double minP[3], minQ[3], mindd;
  int shown_disjoint = 0;
  mindd = v_dist2(&S[0], &T[0]) + 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      orc_seg_points(VEC, P, Q, &S[3 * i], &Sv[3 * i], &T[3 * j], &Tv[3 * j]);
      v_sub(V, Q, P);
      double dd = v_dot(V, V);
      if (dd <= mindd)
      {
        v_cpy(minP, P); v_cpy(minQ, Q); mindd = dd;
        *f1_type = 1; *f1_fid = i; *f2_type = 1; *f2_fid = j;
        v_sub(Z, &S[3 * ((i + 2) % 3)], P);
        double a = v_dot(Z, VEC);
        v_sub(Z, &T[3 * ((j + 2) % 3)], Q);
        double b = v_dot(Z, VEC);
        if ((a <= 0) && (b >= 0)) return sqrt(dd);
        double p = v_dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
    }
  double onFace[3], vert[3];
  {
    // face of S against the vertices of T, :262-339 (the feature ids need the chosen vertex: redo the choice)
    double n[3], nl, proj[3];
    v_cross(n, &Sv[0], &Sv[3]); nl = v_dot(n, n);
    if (nl > 1e-15)
    {
      v_sub(V, &S[0], &T[0]); proj[0] = v_dot(V, n);
      v_sub(V, &S[0], &T[3]); proj[1] = v_dot(V, n);
      v_sub(V, &S[0], &T[6]); proj[2] = v_dot(V, n);
      int point = -1;
      if ((proj[0] > 0) && (proj[1] > 0) && (proj[2] > 0)) { point = (proj[0] < proj[1]) ? 0 : 1; if (proj[2] < proj[point]) point = 2; }
      else if ((proj[0] < 0) && (proj[1] < 0) && (proj[2] < 0)) { point = (proj[0] > proj[1]) ? 0 : 1; if (proj[2] > proj[point]) point = 2; }
      if (face_vertex_case(S, Sv, T, shown_disjoint, onFace, vert))
      {
        *f1_type = 2; *f1_fid = -1; *f2_type = 0; *f2_fid = point;
        v_cpy(P, onFace); v_cpy(Q, vert);
        return sqrt(v_dist2(P, Q));
      }
    }
  }
  {
    double n[3], nl, proj[3];
    v_cross(n, &Tv[0], &Tv[3]); nl = v_dot(n, n);
    if (nl > 1e-15)
    {
      v_sub(V, &T[0], &S[0]); proj[0] = v_dot(V, n);
      v_sub(V, &T[0], &S[3]); proj[1] = v_dot(V, n);
      v_sub(V, &T[0], &S[6]); proj[2] = v_dot(V, n);
      int point = -1;
      if ((proj[0] > 0) && (proj[1] > 0) && (proj[2] > 0)) { point = (proj[0] < proj[1]) ? 0 : 1; if (proj[2] < proj[point]) point = 2; }
      else if ((proj[0] < 0) && (proj[1] < 0) && (proj[2] < 0)) { point = (proj[0] > proj[1]) ? 0 : 1; if (proj[2] > proj[point]) point = 2; }
      if (face_vertex_case(T, Tv, S, shown_disjoint, onFace, vert))
      {
        *f1_type = 0; *f1_fid = point; *f2_type = 2; *f2_fid = -1;
        v_cpy(P, vert); v_cpy(Q, onFace);
        return sqrt(v_dist2(P, Q));
      }
    }
  }
  if (shown_disjoint) { v_cpy(P, minP); v_cpy(Q, minQ); return sqrt(mindd); }
  *collided = 1;
  return 0;
}

namespace {
struct ContactCtx
{
  const orc_bvh *A, *B;
  const int32_t *va, *vb;
  double Rrel[9], Trel[3], threshold;
  int64_t max_out, count;
  orc_contact *out;
};

inline void feature_ids(int32_t out[3], int type, int fid, const int32_t *v)
{
  out[0] = out[1] = out[2] = -1;
  if (!v) return;
  if (type == 0) out[0] = v[fid];
  else if (type == 1) { out[0] = v[fid]; out[1] = v[(fid + 1) % 3]; }
  else if (type == 2) { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; }
}

// TOCStepRecurse_Dis_contact, C2A/src/C2A.cpp:1523-1725
void contact_recurse(ContactCtx &cx, const double R[9], const double T[3], int b1, int b2)
{
  const orc_bvh *A = cx.A, *B = cx.B;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  if (l1 && l2)
  {
    const int ta = -A->first_child[b1] - 1, tb = -B->first_child[b2] - 1;
    double p[3], q[3], tri2[9];
    m_v_p(&tri2[0], cx.Rrel, &B->tris[9 * tb + 0], cx.Trel);
    m_v_p(&tri2[3], cx.Rrel, &B->tris[9 * tb + 3], cx.Trel);
    m_v_p(&tri2[6], cx.Rrel, &B->tris[9 * tb + 6], cx.Trel);
    int32_t f1t = -1, f1f = 0, f2t = -1, f2f = 0, col = 0;
    const double d = orc_tri_dist_features(p, q, &A->tris[9 * ta], tri2, &f1t, &f1f, &f2t, &f2f, &col);
    if (f1t == -1 || f2t == -1) return;
    if (d <= cx.threshold)
    {
      if (cx.count < cx.max_out)
      {
        orc_contact &c = cx.out[cx.count];
        c.type_a = f1t + 1; c.type_b = f2t + 1;
        feature_ids(c.fid_a, f1t, f1f, cx.va ? cx.va + 3 * ta : 0);
        feature_ids(c.fid_b, f2t, f2f, cx.vb ? cx.vb + 3 * tb : 0);
        c.tri_a = ta; c.tri_b = tb;
        v_cpy(c.pa, p);
        double tmp[3];
        v_sub(tmp, q, cx.Trel);
        mt_v(c.pb, cx.Rrel, tmp);
        c.dist = d;
      }
      cx.count++;
    }
    return;
  }
  int a1, a2, c1, c2;
  double R1[9], T1[3], R2[9], T2[3], Tt[3], S[3];
  const double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    a1 = A->first_child[b1]; a2 = b2; c1 = a1 + 1; c2 = b2;
    mt_m(R1, &A->R[9 * a1], R); v_sub(Tt, T, &A->Tr[3 * a1]); mt_v(T1, &A->R[9 * a1], Tt);
    mt_m(R2, &A->R[9 * c1], R); v_sub(Tt, T, &A->Tr[3 * c1]); mt_v(T2, &A->R[9 * c1], Tt);
  }
  else
  {
    a1 = b1; a2 = B->first_child[b2]; c1 = b1; c2 = a2 + 1;
    m_m(R1, R, &B->R[9 * a2]); m_v_p(T1, R, &B->Tr[3 * a2], T);
    m_m(R2, R, &B->R[9 * c2]); m_v_p(T2, R, &B->Tr[3 * c2], T);
  }
  const double d1 = bv_distance(R1, T1, A, a1, B, a2, S);
  const double d2 = bv_distance(R2, T2, A, c1, B, c2, S);
  // res->distance = threshold, abs_err = rel_err = 0 during this pass (:1755-1760)
#define NEAR_ENOUGH(d) (((d) < (cx.threshold - 0)) || ((d) * (1.0 + 0) < cx.threshold))
  if (d2 < d1)
  {
    if (NEAR_ENOUGH(d2)) contact_recurse(cx, R2, T2, c1, c2);
    if (NEAR_ENOUGH(d1)) contact_recurse(cx, R1, T1, a1, a2);
  }
  else
  {
    if (NEAR_ENOUGH(d1)) contact_recurse(cx, R1, T1, a1, a2);
    if (NEAR_ENOUGH(d2)) contact_recurse(cx, R2, T2, c1, c2);
  }
#undef NEAR_ENOUGH
}
}  // namespace

// C2A_QueryContact -> C2A_TimeOfContactStep_Contact, C2A/src/C2A.cpp:1937-1966, 1727-1775
extern "C" int64_t orc_contacts(const orc_bvh *A, const orc_bvh *B, const int32_t *vidx_a, const int32_t *vidx_b,
                                const double pose1[12], const double pose2[12], double threshold, int64_t max_out,
                                orc_contact *out)
{
  ContactCtx cx;
  cx.A = A; cx.B = B; cx.va = vidx_a; cx.vb = vidx_b; cx.threshold = threshold; cx.max_out = max_out; cx.count = 0; cx.out = out;
  const double *R1 = pose1, *T1 = pose1 + 9, *R2 = pose2, *T2 = pose2 + 9;
  double Tt[3], Rt[9], R[9], T[3];
  mt_m(cx.Rrel, R1, R2);
  v_sub(Tt, T2, T1);
  mt_v(cx.Trel, R1, Tt);
  m_m(Rt, cx.Rrel, &B->R[0]);
  mt_m(R, &A->R[0], Rt);
  m_v_p(Tt, cx.Rrel, &B->Tr[0], cx.Trel);
  v_sub(Tt, Tt, &A->Tr[0]);
  mt_v(T, &A->R[0], Tt);
  contact_recurse(cx, R, T, 0, 0);
  return cx.count;
}

// ---- discrete distance query ---------------------------------------------------------------
// C2A_Distance (C2A/src/C2A_PQP.cpp:970-1056) with the depth-first routine it takes for qsize <= 2,
// C2ADistanceRecurse (:481-614).  Seeds stand in for o1->last_tri / o2->last_tri.
namespace {
struct DistCtx
{
  const orc_bvh *A, *B;
  double Rrel[9], Trel[3];
  double rel_err, abs_err;
  orc_distance_result *res;
};

void distance_recurse(DistCtx &cx, const double R[9], const double T[3], int b1, int b2)
{
  const orc_bvh *A = cx.A, *B = cx.B;
  const double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  orc_distance_result *res = cx.res;
  if (l1 && l2)
  {
    res->num_tri_tests++;
    double p[3], q[3];
    const int ta = -A->first_child[b1] - 1, tb = -B->first_child[b2] - 1;
    const double d = orc_tri_distance(cx.Rrel, cx.Trel, &A->tris[9 * ta], &B->tris[9 * tb], p, q);
    if (d < res->distance)
    {
      res->distance = d;
      res->tri_a = ta; res->tri_b = tb;
      v_cpy(res->p1, p); v_cpy(res->p2, q);
    }
    return;
  }
  int a1, a2, c1, c2;
  double R1[9], T1[3], R2[9], T2[3], Tt[3];
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    a1 = A->first_child[b1]; a2 = b2; c1 = a1 + 1; c2 = b2;
    mt_m(R1, &A->R[9 * a1], R); v_sub(Tt, T, &A->Tr[3 * a1]); mt_v(T1, &A->R[9 * a1], Tt);
    mt_m(R2, &A->R[9 * c1], R); v_sub(Tt, T, &A->Tr[3 * c1]); mt_v(T2, &A->R[9 * c1], Tt);
  }
  else
  {
    a1 = b1; a2 = B->first_child[b2]; c1 = b1; c2 = a2 + 1;
    m_m(R1, R, &B->R[9 * a2]); m_v_p(T1, R, &B->Tr[3 * a2], T);
    m_m(R2, R, &B->R[9 * c2]); m_v_p(T2, R, &B->Tr[3 * c2], T);
  }
  res->num_bv_tests += 2;
  double S[3];
  const double d1 = bv_distance(R1, T1, A, a1, B, a2, S);
  const double d2 = bv_distance(R2, T2, A, c1, B, c2, S);
#define CLOSER(d) (((d) < (res->distance - cx.abs_err)) || ((d) * (1 + cx.rel_err) < res->distance))
  if (d2 < d1)
  {
    if (CLOSER(d2)) distance_recurse(cx, R2, T2, c1, c2);
    if (CLOSER(d1)) distance_recurse(cx, R1, T1, a1, a2);
  }
  else
  {
    if (CLOSER(d1)) distance_recurse(cx, R1, T1, a1, a2);
    if (CLOSER(d2)) distance_recurse(cx, R2, T2, c1, c2);
  }
#undef CLOSER
}
}  // namespace

extern "C" void orc_distance(const orc_bvh *A, const orc_bvh *B, const double pose24[24], int32_t seedA, int32_t seedB,
                             double rel_err, double abs_err, orc_distance_result *res)
{
  DistCtx cx;
  cx.A = A; cx.B = B; cx.rel_err = rel_err; cx.abs_err = abs_err; cx.res = res;
  const double *R1 = &pose24[0], *T1 = &pose24[9], *R2 = &pose24[12], *T2 = &pose24[21];
  double Tt[3], Rt[9], R[9], T[3], p[3], q[3];
  mt_m(cx.Rrel, R1, R2);
  v_sub(Tt, T2, T1);
  mt_v(cx.Trel, R1, Tt);
  res->distance = orc_tri_distance(cx.Rrel, cx.Trel, &A->tris[9 * seedA], &B->tris[9 * seedB], p, q);
  res->tri_a = seedA; res->tri_b = seedB;
  v_cpy(res->p1, p); v_cpy(res->p2, q);
  res->num_bv_tests = 0; res->num_tri_tests = 0;
  m_m(Rt, cx.Rrel, &B->R[0]);
  mt_m(R, &A->R[0], Rt);
  m_v_p(Tt, cx.Rrel, &B->Tr[0], cx.Trel);
  v_sub(Tt, Tt, &A->Tr[0]);
  mt_v(T, &A->R[0], Tt);
  distance_recurse(cx, R, T, 0, 0);
  // res->p2 is in cs 1; transform it to cs 2 (:1046-1048)
  double u[3];
  v_sub(u, res->p2, cx.Trel);
  mt_v(res->p2, cx.Rrel, u);
}

// C2A_Distance with qsize > 2: C2ADistanceQueueRecurse (C2A/src/C2A_PQP.cpp:624-787), best-first over a bounded queue of
// pending node pairs, recursing with a fresh queue whenever the current one cannot take two more.  The queue is PQP's BVTQ
// (un-vendored); what matters of it here is only which of several equally distant pairs comes out first: the stand-in the
// compiled reference links (oracle/pqp_shim/BVTQ.h) takes the one inserted earliest, and so does this.
namespace {
struct QTest { double d; int b1, b2; double R[9], T[3]; };

void distance_queue_recurse(DistCtx &cx, int qsize, const QTest &root)
{
  const orc_bvh *A = cx.A, *B = cx.B;
  orc_distance_result *res = cx.res;
  std::vector<QTest> q;
  q.reserve(qsize);
  QTest cur = root;
  while (true)
  {
    const int l1 = A->first_child[cur.b1] < 0, l2 = B->first_child[cur.b2] < 0;
    if (l1 && l2)
    {
      res->num_tri_tests++;
      double p[3], qq[3];
      const int ta = -A->first_child[cur.b1] - 1, tb = -B->first_child[cur.b2] - 1;
      const double d = orc_tri_distance(cx.Rrel, cx.Trel, &A->tris[9 * ta], &B->tris[9 * tb], p, qq);
      if (d < res->distance)
      {
        res->distance = d;
        res->tri_a = ta; res->tri_b = tb;
        v_cpy(res->p1, p); v_cpy(res->p2, qq);
      }
    }
    else if ((int)q.size() == qsize - 1) distance_queue_recurse(cx, qsize, cur);
    else
    {
      const double sz1 = bv_size(A, cur.b1), sz2 = bv_size(B, cur.b2);
      res->num_bv_tests += 2;
      QTest t1, t2;
      double Tt[3], S[3];
      if (l2 || (!l1 && (sz1 > sz2)))
      {
        const int c1 = A->first_child[cur.b1], c2 = c1 + 1;
        t1.b1 = c1; t1.b2 = cur.b2;
        mt_m(t1.R, &A->R[9 * c1], cur.R); v_sub(Tt, cur.T, &A->Tr[3 * c1]); mt_v(t1.T, &A->R[9 * c1], Tt);
        t1.d = bv_distance(t1.R, t1.T, A, t1.b1, B, t1.b2, S);
        t2.b1 = c2; t2.b2 = cur.b2;
        mt_m(t2.R, &A->R[9 * c2], cur.R); v_sub(Tt, cur.T, &A->Tr[3 * c2]); mt_v(t2.T, &A->R[9 * c2], Tt);
        t2.d = bv_distance(t2.R, t2.T, A, t2.b1, B, t2.b2, S);
      }
      else
      {
        const int c1 = B->first_child[cur.b2], c2 = c1 + 1;
        t1.b1 = cur.b1; t1.b2 = c1;
        m_m(t1.R, cur.R, &B->R[9 * c1]); m_v_p(t1.T, cur.R, &B->Tr[3 * c1], cur.T);
        t1.d = bv_distance(t1.R, t1.T, A, t1.b1, B, t1.b2, S);
        t2.b1 = cur.b1; t2.b2 = c2;
        m_m(t2.R, cur.R, &B->R[9 * c2]); m_v_p(t2.T, cur.R, &B->Tr[3 * c2], cur.T);
        t2.d = bv_distance(t2.R, t2.T, A, t2.b1, B, t2.b2, S);
      }
      q.push_back(t1);
      q.push_back(t2);
    }
    if (q.empty()) break;
    size_t k = 0;
    for (size_t i = 1; i < q.size(); i++) if (q[i].d < q[k].d) k = i;
    cur = q[k];
    q.erase(q.begin() + k);
    if ((cur.d + cx.abs_err >= res->distance) && ((cur.d * (1 + cx.rel_err)) >= res->distance)) break;
  }
}
}  // namespace

extern "C" void orc_distance_queue(const orc_bvh *A, const orc_bvh *B, const double pose24[24], int32_t seedA, int32_t seedB,
                                   double rel_err, double abs_err, int32_t qsize, orc_distance_result *res)
{
  if (qsize <= 2) { orc_distance(A, B, pose24, seedA, seedB, rel_err, abs_err, res); return; }   // :1031-1034
  DistCtx cx;
  cx.A = A; cx.B = B; cx.rel_err = rel_err; cx.abs_err = abs_err; cx.res = res;
  const double *R1 = &pose24[0], *T1 = &pose24[9], *R2 = &pose24[12], *T2 = &pose24[21];
  double Tt[3], Rt[9], p[3], q[3];
  mt_m(cx.Rrel, R1, R2);
  v_sub(Tt, T2, T1);
  mt_v(cx.Trel, R1, Tt);
  res->distance = orc_tri_distance(cx.Rrel, cx.Trel, &A->tris[9 * seedA], &B->tris[9 * seedB], p, q);
  res->tri_a = seedA; res->tri_b = seedB;
  v_cpy(res->p1, p); v_cpy(res->p2, q);
  res->num_bv_tests = 0; res->num_tri_tests = 0;
  QTest root;
  root.b1 = 0; root.b2 = 0; root.d = 0;
  m_m(Rt, cx.Rrel, &B->R[0]);
  mt_m(root.R, &A->R[0], Rt);
  m_v_p(Tt, cx.Rrel, &B->Tr[0], cx.Trel);
  v_sub(Tt, Tt, &A->Tr[0]);
  mt_v(root.T, &A->R[0], Tt);
  distance_queue_recurse(cx, qsize, root);
  double u[3];
  v_sub(u, res->p2, cx.Trel);
  mt_v(res->p2, cx.Rrel, u);
}

// ---- discrete collision queries -----------------------------------------------------------------
// C2A_Collide, both overloads (C2A/src/C2A_PQP.cpp:798-968 and :1060-1280).  Their box-overlap and triangle-overlap
// tests are PQP's (obb_disjoint, TriContact: un-vendored, no in-tree source): restated here from the published
// separating-axis formulations, in the arithmetic of oracle/pqp_shim/pqp_shim.cpp, which is what the compiled
// reference (_ref) links -- parity for these two functions is therefore pinned to the shim, not to PQP itself.
extern "C" int32_t orc_obb_disjoint(const double B[9], const double T[3], const double a[3], const double b[3])
{
  const double reps = 1e-6;
  double Bf[9];
  for (int i = 0; i < 9; i++) Bf[i] = fabs(B[i]) + reps;
  for (int i = 0; i < 3; i++)
  {
    const double t = fabs(T[i]);
    if (t > a[i] + b[0] * Bf[3 * i + 0] + b[1] * Bf[3 * i + 1] + b[2] * Bf[3 * i + 2]) return 1 + i;
  }
  for (int j = 0; j < 3; j++)
  {
    const double s = T[0] * B[0 + j] + T[1] * B[3 + j] + T[2] * B[6 + j];
    const double t = fabs(s);
    if (t > b[j] + a[0] * Bf[0 + j] + a[1] * Bf[3 + j] + a[2] * Bf[6 + j]) return 4 + j;
  }
  int code = 7;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++, code++)
    {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double s = T[i2] * B[3 * i1 + j] - T[i1] * B[3 * i2 + j];
      const double t = fabs(s);
      const double ra = a[i1] * Bf[3 * i2 + j] + a[i2] * Bf[3 * i1 + j];
      const double rb = b[j1] * Bf[3 * i + j2] + b[j2] * Bf[3 * i + j1];
      if (t > ra + rb) return code;
    }
  return 0;
}

namespace {
inline void cross3(double r[3], const double a[3], const double b[3])
{
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
inline bool axis_separates(const double ax[3], const double p[3][3], const double q[3][3])
{
  double pmin = v_dot(ax, p[0]), pmax = pmin, qmin = v_dot(ax, q[0]), qmax = qmin;
  for (int i = 1; i < 3; i++)
  {
    const double v = v_dot(ax, p[i]); if (v < pmin) pmin = v; if (v > pmax) pmax = v;
    const double w = v_dot(ax, q[i]); if (w < qmin) qmin = w; if (w > qmax) qmax = w;
  }
  return (pmin > qmax) || (qmin > pmax);
}
}  // namespace

// P: triangle 1 (p1,p2,p3), Q: triangle 2 already in triangle 1's frame
extern "C" int32_t orc_tri_contact(const double P[9], const double Q[9])
{
  double p[3][3], q[3][3], e[3][3], f[3][3], n[3], m[3], ax[3];
  for (int k = 0; k < 3; k++) { v_sub(p[k], &P[3 * k], &P[0]); v_sub(q[k], &Q[3 * k], &P[0]); }
  v_sub(e[0], p[1], p[0]); v_sub(e[1], p[2], p[1]); v_sub(e[2], p[0], p[2]);
  v_sub(f[0], q[1], q[0]); v_sub(f[1], q[2], q[1]); v_sub(f[2], q[0], q[2]);
  cross3(n, e[0], e[1]);
  cross3(m, f[0], f[1]);
  if (axis_separates(n, p, q)) return 0;
  if (axis_separates(m, p, q)) return 0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      cross3(ax, e[i], f[j]);
      if (axis_separates(ax, p, q)) return 0;
    }
  for (int i = 0; i < 3; i++)
  {
    cross3(ax, e[i], n); if (axis_separates(ax, p, q)) return 0;
    cross3(ax, f[i], m); if (axis_separates(ax, p, q)) return 0;
  }
  return 1;
}

namespace {
struct CollideCtx
{
  const orc_bvh *A, *B;
  const double *dA, *ToA, *dB, *ToB;  // OBB half-dimensions and centres, [n_nodes][3]
  double Rrel[9], Trel[3];
  int flag, max_pairs, num_pairs, num_bv_tests, num_tri_tests;
  int32_t *pairs;
};

// CollideRecurse, C2A_PQP.cpp:798-909 (transforms chained through the box centres To, OBB_TYPE branch)
void collide_recurse(CollideCtx &cx, const double R[9], const double T[3], int b1, int b2)
{
  const orc_bvh *A = cx.A, *B = cx.B;
  cx.num_bv_tests++;
  if (orc_obb_disjoint(R, T, &cx.dA[3 * b1], &cx.dB[3 * b2]) != 0) return;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  if (l1 && l2)
  {
    cx.num_tri_tests++;
    const int ta = -A->first_child[b1] - 1, tb = -B->first_child[b2] - 1;
    double Q[9];
    for (int k = 0; k < 3; k++) m_v_p(&Q[3 * k], cx.Rrel, &B->tris[9 * tb + 3 * k], cx.Trel);
    if (orc_tri_contact(&A->tris[9 * ta], Q))
    {
      if (cx.num_pairs < cx.max_pairs) { cx.pairs[2 * cx.num_pairs] = ta; cx.pairs[2 * cx.num_pairs + 1] = tb; }
      cx.num_pairs++;
    }
    return;
  }
  const double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  double Rc[9], Tc[3], Tt[3];
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    const int c1 = A->first_child[b1], c2 = c1 + 1;
    mt_m(Rc, &A->R[9 * c1], R); v_sub(Tt, T, &cx.ToA[3 * c1]); mt_v(Tc, &A->R[9 * c1], Tt);
    collide_recurse(cx, Rc, Tc, c1, b2);
    if (cx.flag == 2 && cx.num_pairs > 0) return;
    mt_m(Rc, &A->R[9 * c2], R); v_sub(Tt, T, &cx.ToA[3 * c2]); mt_v(Tc, &A->R[9 * c2], Tt);
    collide_recurse(cx, Rc, Tc, c2, b2);
  }
  else
  {
    const int c1 = B->first_child[b2], c2 = c1 + 1;
    m_m(Rc, R, &B->R[9 * c1]); m_v_p(Tc, R, &cx.ToB[3 * c1], T);
    collide_recurse(cx, Rc, Tc, b1, c1);
    if (cx.flag == 2 && cx.num_pairs > 0) return;
    m_m(Rc, R, &B->R[9 * c2]); m_v_p(Tc, R, &cx.ToB[3 * c2], T);
    collide_recurse(cx, Rc, Tc, b1, c2);
  }
}
}  // namespace

// C2A_Collide(PQP_CollideResult*, ..., flag), C2A_PQP.cpp:910-968.  flag 1: all contacts, 2: first contact.  pairs: builder-order
// triangle indices (the reference reports Tri::id) of the first min(num_pairs, max_pairs) pairs in traversal order.
extern "C" int32_t orc_collide(const orc_bvh *A, const orc_bvh *B, const double *dA, const double *ToA, const double *dB,
                               const double *ToB, const double pose24[24], int32_t flag, int32_t max_pairs, int32_t *pairs,
                               int32_t *num_bv_tests, int32_t *num_tri_tests)
{
  CollideCtx cx;
  cx.A = A; cx.B = B; cx.dA = dA; cx.ToA = ToA; cx.dB = dB; cx.ToB = ToB;
  cx.flag = flag; cx.max_pairs = max_pairs; cx.pairs = pairs; cx.num_pairs = cx.num_bv_tests = cx.num_tri_tests = 0;
  const double *R1 = &pose24[0], *T1 = &pose24[9], *R2 = &pose24[12], *T2 = &pose24[21];
  double Tt[3], Rt[9], R[9], T[3];
  mt_m(cx.Rrel, R1, R2);
  v_sub(Tt, T2, T1);
  mt_v(cx.Trel, R1, Tt);
  m_m(Rt, cx.Rrel, &B->R[0]);
  mt_m(R, &A->R[0], Rt);
  m_v_p(Tt, cx.Rrel, &ToB[0], cx.Trel);
  v_sub(Tt, Tt, &ToA[0]);
  mt_v(T, &A->R[0], Tt);
  collide_recurse(cx, R, T, 0, 0);
  if (num_bv_tests) *num_bv_tests = cx.num_bv_tests;
  if (num_tri_tests) *num_tri_tests = cx.num_tri_tests;
  return cx.num_pairs;
}

namespace {
struct CollideDistCtx { DistCtx base; const double *dA, *dB; };

// C2ACollideRecurse, C2A_PQP.cpp:1060-1196: C2ADistanceRecurse behind a box-overlap gate.  The gate is handed the
// transform chained through the RSS corners Tr (the RSS_TYPE branch comes first in the source), as the reference does.
void collide_distance_recurse(CollideDistCtx &cx, const double R[9], const double T[3], int b1, int b2)
{
  const orc_bvh *A = cx.base.A, *B = cx.base.B;
  const double sz1 = bv_size(A, b1), sz2 = bv_size(B, b2);
  if (orc_obb_disjoint(R, T, &cx.dA[3 * b1], &cx.dB[3 * b2]) != 0) return;
  const int l1 = A->first_child[b1] < 0, l2 = B->first_child[b2] < 0;
  orc_distance_result *res = cx.base.res;
  if (l1 && l2)
  {
    res->num_tri_tests++;
    double p[3], q[3];
    const int ta = -A->first_child[b1] - 1, tb = -B->first_child[b2] - 1;
    const double d = orc_tri_distance(cx.base.Rrel, cx.base.Trel, &A->tris[9 * ta], &B->tris[9 * tb], p, q);
    if (d < res->distance)
    {
      res->distance = d;
      res->tri_a = ta; res->tri_b = tb;
      v_cpy(res->p1, p); v_cpy(res->p2, q);
    }
    return;
  }
  int a1, a2, c1, c2;
  double R1[9], T1[3], R2[9], T2[3], Tt[3];
  if (l2 || (!l1 && (sz1 > sz2)))
  {
    a1 = A->first_child[b1]; a2 = b2; c1 = a1 + 1; c2 = b2;
    mt_m(R1, &A->R[9 * a1], R); v_sub(Tt, T, &A->Tr[3 * a1]); mt_v(T1, &A->R[9 * a1], Tt);
    mt_m(R2, &A->R[9 * c1], R); v_sub(Tt, T, &A->Tr[3 * c1]); mt_v(T2, &A->R[9 * c1], Tt);
  }
  else
  {
    a1 = b1; a2 = B->first_child[b2]; c1 = b1; c2 = a2 + 1;
    m_m(R1, R, &B->R[9 * a2]); m_v_p(T1, R, &B->Tr[3 * a2], T);
    m_m(R2, R, &B->R[9 * c2]); m_v_p(T2, R, &B->Tr[3 * c2], T);
  }
  res->num_bv_tests += 2;
  double S[3];
  const double d1 = bv_distance(R1, T1, A, a1, B, a2, S);
  const double d2 = bv_distance(R2, T2, A, c1, B, c2, S);
#define CLOSER(d) (((d) < (res->distance - cx.base.abs_err)) || ((d) * (1 + cx.base.rel_err) < res->distance))
  if (d2 < d1)
  {
    if (CLOSER(d2)) collide_distance_recurse(cx, R2, T2, c1, c2);
    if (CLOSER(d1)) collide_distance_recurse(cx, R1, T1, a1, a2);
  }
  else
  {
    if (CLOSER(d1)) collide_distance_recurse(cx, R1, T1, a1, a2);
    if (CLOSER(d2)) collide_distance_recurse(cx, R2, T2, c1, c2);
  }
#undef CLOSER
}
}  // namespace

// C2A_Collide(C2A_DistanceResult*, ..., rel_err, abs_err), C2A_PQP.cpp:1199-1280
extern "C" void orc_collide_distance(const orc_bvh *A, const orc_bvh *B, const double *dA, const double *dB, const double pose24[24],
                                     int32_t seedA, int32_t seedB, double rel_err, double abs_err, orc_distance_result *res)
{
  CollideDistCtx cx;
  cx.base.A = A; cx.base.B = B; cx.base.rel_err = rel_err; cx.base.abs_err = abs_err; cx.base.res = res; cx.dA = dA; cx.dB = dB;
  const double *R1 = &pose24[0], *T1 = &pose24[9], *R2 = &pose24[12], *T2 = &pose24[21];
  double Tt[3], Rt[9], R[9], T[3], p[3], q[3];
  mt_m(cx.base.Rrel, R1, R2);
  v_sub(Tt, T2, T1);
  mt_v(cx.base.Trel, R1, Tt);
  res->distance = orc_tri_distance(cx.base.Rrel, cx.base.Trel, &A->tris[9 * seedA], &B->tris[9 * seedB], p, q);
  res->tri_a = seedA; res->tri_b = seedB;
  v_cpy(res->p1, p); v_cpy(res->p2, q);
  res->num_bv_tests = 0; res->num_tri_tests = 0;
  m_m(Rt, cx.base.Rrel, &B->R[0]);
  mt_m(R, &A->R[0], Rt);
  m_v_p(Tt, cx.base.Rrel, &B->Tr[0], cx.base.Trel);
  v_sub(Tt, Tt, &A->Tr[0]);
  mt_v(T, &A->R[0], Tt);
  collide_distance_recurse(cx, R, T, 0, 0);
  double u[3];
  v_sub(u, res->p2, cx.base.Trel);
  mt_v(res->p2, cx.base.Rrel, u);
}

// Design study entry: orc_solve with every exact-mode CA step run through the speculative subtree split at
// frontier depth K; results must equal orc_solve's bit for bit (tests/test_oracle.py), stats accumulate.
extern "C" void orc_solve_spec(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB,
                               double tol_d, double tol_t, int32_t K, orc_result *out, orc_spec_stats *stats)
{
  stats->depth = K;
  g_spec = stats; g_spec_prev = -1;
  orc_solve(A, B, poses, seedA, seedB, tol_d, tol_t, out);
  g_spec = nullptr;
}

// Design study entry: orc_solve that also returns the sequence of node pairs the traversal visits (expanded pairs and
// leaf pairs as (b1 << 32 | b2), a ~0 word before each CA step).  Returns the number of words; at most cap are written.
extern "C" int64_t orc_solve_visits(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB,
                                    double tol_d, double tol_t, orc_result *out, uint64_t *visits, int64_t cap)
{
  std::vector<uint64_t> v;
  g_visits = &v;
  orc_solve(A, B, poses, seedA, seedB, tol_d, tol_t, out);
  g_visits = nullptr;
  const int64_t n = (int64_t)v.size();
  if (visits) memcpy(visits, v.data(), sizeof(uint64_t) * (size_t)(n < cap ? n : cap));
  return n;
}

// Round-2 algorithm entry: orc_solve with every exact-mode CA step run as prefix + wide phase (see wide_step); results must
// equal orc_solve's bit for bit.  stats->window (pairs popped per round) is an input (0 -> 16).
extern "C" void orc_solve_wide(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB,
                               double tol_d, double tol_t, orc_result *out, orc_wide_stats *stats)
{
  if (stats->window <= 0) stats->window = 16;
  if (stats->leaf_batch < 0) stats->leaf_batch = 0;
  g_wide = stats;
  orc_solve(A, B, poses, seedA, seedB, tol_d, tol_t, out);
  g_wide = nullptr;
}

// Design study entry: orc_solve that also reports the CA-loop state at the moment numCA reaches `step` (lamda, distance, mint,
// BV and triangle tests so far; zeros if the query ends earlier): how well does the state after a few cheap steps predict
// the cost of the whole query?
extern "C" void orc_solve_probe(const orc_bvh *A, const orc_bvh *B, const double *poses, int64_t n, double tol_d, double tol_t,
                                int32_t step, orc_result *out, double *probe5, int32_t n_threads)
{
  std::atomic<int64_t> next(0);
  auto work = [&](int) {
    while (true)
    {
      const int64_t i = next.fetch_add(1);
      if (i >= n) break;
      for (int k = 0; k < 5; k++) probe5[5 * i + k] = 0;
      g_probe = probe5 + 5 * i; g_probe_step = step;
      orc_solve(A, B, poses + 48 * i, 0, 0, tol_d, tol_t, &out[i]);
      g_probe = nullptr;
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < (n_threads < 1 ? 1 : n_threads); t++) th.emplace_back(work, t);
  for (auto &t : th) t.join();
}

// Design study entry: orc_solve with every CA step replayed over the previous step's visit list (see replay_step).
extern "C" void orc_solve_replay(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB,
                                 double tol_d, double tol_t, orc_result *out, orc_replay_stats *stats)
{
  g_rp = stats;
  rp_reset();
  orc_solve(A, B, poses, seedA, seedB, tol_d, tol_t, out);
  rp_reset();
  g_rp = nullptr;
}
