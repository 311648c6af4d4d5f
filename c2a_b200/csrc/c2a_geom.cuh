// Device geometry for the C2A CCD hot path (sm_100a, FP64, compiled with -fmad=false).
//
// Every arithmetic expression keeps the operand order of the reference (paths relative
// to /root/reference) so that results are bit-identical to its CPU build compiled with
// -ffp-contract=off; only control flow is reshaped for SIMT execution.
//
//   rss_rect_dist   <- C2ARectDist            C2A/C2A_RectDist.h:157-934
//   seg_points      <- PQP SegPoints          (in-tree copy C2A/src/C2A.cpp:59-163)
//   tri_dist        <- PQP TriDist            (in-tree copy C2A/src/C2A.cpp:165-405)
//   tri_distance    <- PQP TriDistance        (call sites C2A/src/C2A.cpp:1148,1916)
#pragma once
#include <cuda_runtime.h>

namespace c2a {

#define C2A_DEV __device__ __forceinline__

// ---- small fixed-size linear algebra (row-major 3x3), PQP MatVec.h conventions ----
C2A_DEV void v_sub(double r[3], const double a[3], const double b[3]) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
C2A_DEV void v_add(double r[3], const double a[3], const double b[3]) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
C2A_DEV void v_cpy(double r[3], const double a[3]) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
C2A_DEV void v_madd(double r[3], const double a[3], const double b[3], double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
C2A_DEV double v_dot(const double a[3], const double b[3]) { return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]); }
C2A_DEV void v_cross(double r[3], const double a[3], const double b[3])
{
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
C2A_DEV double v_dist2(const double a[3], const double b[3])
{
  return ((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
}
C2A_DEV double v_len(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
C2A_DEV void v_normalize(double a[3])
{
  double d = 1.0 / sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  a[0] *= d; a[1] *= d; a[2] *= d;
}
C2A_DEV void m_v(double r[3], const double M[9], const double v[3])
{
  r[0] = (M[0] * v[0] + M[1] * v[1] + M[2] * v[2]);
  r[1] = (M[3] * v[0] + M[4] * v[1] + M[5] * v[2]);
  r[2] = (M[6] * v[0] + M[7] * v[1] + M[8] * v[2]);
}
C2A_DEV void m_v_p(double r[3], const double M[9], const double v[3], const double t[3])
{
  r[0] = (M[0] * v[0] + M[1] * v[1] + M[2] * v[2] + t[0]);
  r[1] = (M[3] * v[0] + M[4] * v[1] + M[5] * v[2] + t[1]);
  r[2] = (M[6] * v[0] + M[7] * v[1] + M[8] * v[2] + t[2]);
}
C2A_DEV void mt_v(double r[3], const double M[9], const double v[3])
{
  r[0] = (M[0] * v[0] + M[3] * v[1] + M[6] * v[2]);
  r[1] = (M[1] * v[0] + M[4] * v[1] + M[7] * v[2]);
  r[2] = (M[2] * v[0] + M[5] * v[1] + M[8] * v[2]);
}
C2A_DEV void m_m(double r[9], const double A[9], const double B[9])
{
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      r[3 * i + j] = (A[3 * i + 0] * B[0 + j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]);
}
C2A_DEV void mt_m(double r[9], const double A[9], const double B[9])
{
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      r[3 * i + j] = (A[0 + i] * B[0 + j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j]);
}

// ---- rectangle-rectangle distance ----------------------------------------------------
C2A_DEV void clamp_to(double &v, double lo, double hi) { if (v < lo) v = lo; else if (v > hi) v = hi; }

// C2A/C2A_RectDist.h:81-111
C2A_DEV void seg_params(double &t, double &u, double a, double b, double AdB, double AdT, double BdT)
{
  double denom = 1 - (AdB) * (AdB);
  if (denom == 0) t = 0;
  else { t = (AdT - BdT * AdB) / denom; clamp_to(t, 0, a); }
  u = t * AdB - BdT;
  if (u < 0) { u = 0; t = AdT; clamp_to(t, 0, a); }
  else if (u > b) { u = b; t = u * AdB + AdT; clamp_to(t, 0, a); }
}

// C2A/C2A_RectDist.h:123-154
C2A_DEV bool in_voronoi(double a, double b, double AnB, double AnT, double AdB, double AdT, double BdT)
{
  if (((AnB < 0) ? -AnB : AnB) < 1e-7) return false;
  double t, u, v;
  u = -AnT / AnB; clamp_to(u, 0, b);
  t = u * AdB + AdT; clamp_to(t, 0, a);
  v = t * AdB - BdT;
  if (AnB > 0) return v > (u + 1e-7);
  return v < (u - 1e-7);
}

// Distance between rectangle A ([0,a0]x[0,a1] in its own z=0 plane) and rectangle B placed by
// (R,T) in A's frame; S receives Q-P (closest points) in A's frame.  When both face separations
// are negative S is left untouched, exactly as the reference.
//
// SIMT reshaping: the reference tests 16 edge pairs in a fixed if-ladder.  Here the 16 cheap
// entry predicates (and the trivially-true halves of the acceptance tests) are evaluated first into
// bit masks; each lane then visits only ITS OWN candidate pairs in ladder order, two per trip, forming
// each pair's operands with selects (no 16-way branch) and running the Voronoi tests (the FP64
// divisions) in code common to all lanes.  The expressions and the acceptance order are the
// reference's, so the result is bit-identical.
#ifdef C2A_RD_STATS   // development aid (scripts/build_variant.py ... -DC2A_RD_STATS): candidate / trip histograms
__device__ unsigned long long g_rd_stats[64];
#endif
C2A_DEV double rss_rect_dist(const double R[9], const double T[3], double a0, double a1, double b0, double b1,
                             double S[3])
{
  const double A0B0 = R[0], A0B1 = R[1], A1B0 = R[3], A1B1 = R[4];
  const double aA0B0 = a0 * A0B0, aA0B1 = a0 * A0B1, aA1B0 = a1 * A1B0, aA1B1 = a1 * A1B1;
  const double bA0B0 = b0 * A0B0, bA1B0 = b0 * A1B0, bA0B1 = b1 * A0B1, bA1B1 = b1 * A1B1;
  double Tba[3];
  mt_v(Tba, R, T);

  // extents of A's corners along B's axes and of B's corners along A's axes (:193-745)
  const double ALL_x = -Tba[0], ALU_x = ALL_x + aA1B0, AUL_x = ALL_x + aA0B0, AUU_x = ALU_x + aA0B0;
  const double ALL_y = -Tba[1], ALU_y = ALL_y + aA1B1, AUL_y = ALL_y + aA0B1, AUU_y = ALU_y + aA0B1;
  const double BLL_x = T[0], BLU_x = BLL_x + bA0B1, BUL_x = BLL_x + bA0B0, BUU_x = BLU_x + bA0B0;
  const double BLL_y = T[1], BLU_y = BLL_y + bA1B1, BUL_y = BLL_y + bA1B0, BUU_y = BLU_y + bA1B0;

  bool s;
  s = ALL_x < ALU_x;
  const double LA1_lx = s ? ALL_x : ALU_x, LA1_ux = s ? ALU_x : ALL_x, UA1_lx = s ? AUL_x : AUU_x, UA1_ux = s ? AUU_x : AUL_x;
  s = BLL_x < BLU_x;
  const double LB1_lx = s ? BLL_x : BLU_x, LB1_ux = s ? BLU_x : BLL_x, UB1_lx = s ? BUL_x : BUU_x, UB1_ux = s ? BUU_x : BUL_x;
  s = ALL_y < ALU_y;
  const double LA1_ly = s ? ALL_y : ALU_y, LA1_uy = s ? ALU_y : ALL_y, UA1_ly = s ? AUL_y : AUU_y, UA1_uy = s ? AUU_y : AUL_y;
  s = BLL_x < BUL_x;
  const double LB0_lx = s ? BLL_x : BUL_x, LB0_ux = s ? BUL_x : BLL_x, UB0_lx = s ? BLU_x : BUU_x, UB0_ux = s ? BUU_x : BLU_x;
  s = ALL_x < AUL_x;
  const double LA0_lx = s ? ALL_x : AUL_x, LA0_ux = s ? AUL_x : ALL_x, UA0_lx = s ? ALU_x : AUU_x, UA0_ux = s ? AUU_x : ALU_x;
  s = BLL_y < BLU_y;
  const double LB1_ly = s ? BLL_y : BLU_y, LB1_uy = s ? BLU_y : BLL_y, UB1_ly = s ? BUL_y : BUU_y, UB1_uy = s ? BUU_y : BUL_y;
  s = ALL_y < AUL_y;
  const double LA0_ly = s ? ALL_y : AUL_y, LA0_uy = s ? AUL_y : ALL_y, UA0_ly = s ? ALU_y : AUU_y, UA0_uy = s ? AUU_y : ALU_y;
  s = BLL_y < BUL_y;
  const double LB0_ly = s ? BLL_y : BUL_y, LB0_uy = s ? BUL_y : BLL_y, UB0_ly = s ? BLU_y : BUU_y, UB0_uy = s ? BUU_y : BLU_y;

  // bit k set <=> edge pair k passes its entry predicate.  k = 4*group + sub;
  // group: 0 (A1,B1) 1 (A1,B0) 2 (A0,B1) 3 (A0,B0); sub: 0 (U,U) 1 (U,L) 2 (L,U) 3 (L,L) = the ladder order.
  unsigned mask = 0;
  mask |= ((UA1_ux > b0) && (UB1_ux > a0)) ? 1u << 0 : 0u;
  mask |= ((UA1_lx < 0) && (LB1_ux > a0)) ? 1u << 1 : 0u;
  mask |= ((LA1_ux > b0) && (UB1_lx < 0)) ? 1u << 2 : 0u;
  mask |= ((LA1_lx < 0) && (LB1_lx < 0)) ? 1u << 3 : 0u;
  mask |= ((UA1_uy > b1) && (UB0_ux > a0)) ? 1u << 4 : 0u;
  mask |= ((UA1_ly < 0) && (LB0_ux > a0)) ? 1u << 5 : 0u;
  mask |= ((LA1_uy > b1) && (UB0_lx < 0)) ? 1u << 6 : 0u;
  mask |= ((LA1_ly < 0) && (LB0_lx < 0)) ? 1u << 7 : 0u;
  mask |= ((UA0_ux > b0) && (UB1_uy > a1)) ? 1u << 8 : 0u;
  mask |= ((UA0_lx < 0) && (LB1_uy > a1)) ? 1u << 9 : 0u;
  mask |= ((LA0_ux > b0) && (UB1_ly < 0)) ? 1u << 10 : 0u;
  mask |= ((LA0_lx < 0) && (LB1_ly < 0)) ? 1u << 11 : 0u;
  mask |= ((UA0_uy > b1) && (UB0_uy > a1)) ? 1u << 12 : 0u;
  mask |= ((UA0_ly < 0) && (LB0_uy > a1)) ? 1u << 13 : 0u;
  mask |= ((LA0_uy > b1) && (UB0_ly < 0)) ? 1u << 14 : 0u;
  mask |= ((LA0_ly < 0) && (LB0_ly < 0)) ? 1u << 15 : 0u;

  // trivially-true halves of each pair's acceptance test (the left operand of each || in the ladder)
  unsigned trivA = 0, trivB = 0;
  trivA |= (UA1_lx > b0) ? 1u << 0 : 0u;   trivB |= (UB1_lx > a0) ? 1u << 0 : 0u;
  trivA |= (UA1_ux < 0) ? 1u << 1 : 0u;    trivB |= (LB1_lx > a0) ? 1u << 1 : 0u;
  trivA |= (LA1_lx > b0) ? 1u << 2 : 0u;   trivB |= (UB1_ux < 0) ? 1u << 2 : 0u;
  trivA |= (LA1_ux < 0) ? 1u << 3 : 0u;    trivB |= (LB1_ux < 0) ? 1u << 3 : 0u;
  trivA |= (UA1_ly > b1) ? 1u << 4 : 0u;   trivB |= (UB0_lx > a0) ? 1u << 4 : 0u;
  trivA |= (UA1_uy < 0) ? 1u << 5 : 0u;    trivB |= (LB0_lx > a0) ? 1u << 5 : 0u;
  trivA |= (LA1_ly > b1) ? 1u << 6 : 0u;   trivB |= (UB0_ux < 0) ? 1u << 6 : 0u;
  trivA |= (LA1_uy < 0) ? 1u << 7 : 0u;    trivB |= (LB0_ux < 0) ? 1u << 7 : 0u;
  trivA |= (UA0_lx > b0) ? 1u << 8 : 0u;   trivB |= (UB1_ly > a1) ? 1u << 8 : 0u;
  trivA |= (UA0_ux < 0) ? 1u << 9 : 0u;    trivB |= (LB1_ly > a1) ? 1u << 9 : 0u;
  trivA |= (LA0_lx > b0) ? 1u << 10 : 0u;  trivB |= (UB1_uy < 0) ? 1u << 10 : 0u;
  trivA |= (LA0_ux < 0) ? 1u << 11 : 0u;   trivB |= (LB1_uy < 0) ? 1u << 11 : 0u;
  trivA |= (UA0_ly > b1) ? 1u << 12 : 0u;  trivB |= (UB0_ly > a1) ? 1u << 12 : 0u;
  trivA |= (UA0_uy < 0) ? 1u << 13 : 0u;   trivB |= (LB0_ly > a1) ? 1u << 13 : 0u;
  trivA |= (LA0_ly > b1) ? 1u << 14 : 0u;  trivB |= (UB0_uy < 0) ? 1u << 14 : 0u;
  trivA |= (LA0_uy < 0) ? 1u << 15 : 0u;   trivB |= (LB0_uy < 0) ? 1u << 15 : 0u;

  // Operands of pair k without a 16-way branch.  With ea/eb the edge axes of the pair (group bits) and
  // ua/ub its upper/lower sides, every operand of the ladder is one of a few expressions over scalars
  // chosen by (ea, eb); the handful of pairs whose source spells a sum in another order (which rounds
  // differently) are selected explicitly, so each value is bit-identical to the ladder's.
  struct PairOps { double la, lb, AdB, nA, tA, dA, eA, nB, tB, dB, eB; };
  auto pair_ops = [&](int k, PairOps &o) {
    const int g = k >> 2;
    const bool ea = g < 2, eb = !(g & 1), ua = !(k & 2), ub = !(k & 1);
    o.la = ea ? a1 : a0; o.lb = eb ? b1 : b0;
    const double aO = ea ? a0 : a1, bO = eb ? b0 : b1;
    o.AdB = ea ? (eb ? A1B1 : A1B0) : (eb ? A0B1 : A0B0);                 // R[ea][eb]
    const double Reo = ea ? (eb ? A1B0 : A1B1) : (eb ? A0B0 : A0B1);      // R[ea][ob]
    const double Roe = ea ? (eb ? A0B1 : A0B0) : (eb ? A1B1 : A1B0);      // R[oa][eb]
    const double aRoo = ea ? (eb ? aA0B0 : aA0B1) : (eb ? aA1B0 : aA1B1);  // a[oa]*R[oa][ob]
    const double aRoe = ea ? (eb ? aA0B1 : aA0B0) : (eb ? aA1B1 : aA1B0);  // a[oa]*R[oa][eb]
    const double bReo = ea ? (eb ? bA1B0 : bA1B1) : (eb ? bA0B0 : bA0B1);  // b[ob]*R[ea][ob]
    const double bRoo = ea ? (eb ? bA0B0 : bA0B1) : (eb ? bA1B0 : bA1B1);  // b[ob]*R[oa][ob]
    const double TE = ea ? T[1] : T[0], TO = ea ? T[0] : T[1];
    const double TbaE = eb ? Tba[1] : Tba[0], TbaO = eb ? Tba[0] : Tba[1];
    o.nA = ub ? Reo : -Reo;
    o.nB = ua ? Roe : -Roe;
    const double tA_uu = (g == 0) ? (aRoo - bO - TbaO) : (aRoo - TbaO - bO);
    const double tA_lu = (g == 2) ? (-bO - TbaO) : (-TbaO - bO);
    o.tA = ua ? (ub ? tA_uu : (TbaO - aRoo)) : (ub ? tA_lu : TbaO);
    o.dA = ua ? (aRoe - TbaE) : -TbaE;
    const double eA_u = (g == 2 && !ua) ? (-bReo - TE) : (-TE - bReo);
    o.eA = ub ? eA_u : -TE;
    const double tB_uu = (g == 0) ? (TO + bRoo - aO) : (TO - aO + bRoo);
    o.tB = ua ? (ub ? tB_uu : (TO - aO)) : (ub ? (-TO - bRoo) : -TO);
    o.dB = ub ? (TE + bReo) : TE;
    o.eB = ua ? (TbaE - aRoe) : TbaE;
  };

  // Each lane walks ITS candidate pairs in ladder order, two per trip: the four Voronoi tests of a trip
  // are independent (instruction-level parallelism for the FP64 divisions), and the second pair only
  // counts when the first is rejected, so the accepted pair is the ladder's.
  int kf = -1;
  double t = 0, u = 0;
#ifdef C2A_RD_STATS
  atomicAdd(&g_rd_stats[__popc(mask)], 1ull);
  int rd_trips = 0;
  { const unsigned act = __activemask(); if ((threadIdx.x & 31) == __ffs(act) - 1) { atomicAdd(&g_rd_stats[60], 1ull); atomicAdd(&g_rd_stats[61], (unsigned long long)__popc(act)); } }
#endif
  while (mask)
  {
#ifdef C2A_RD_STATS
    rd_trips++;
    { const unsigned act = __activemask(); if ((threadIdx.x & 31) == __ffs(act) - 1) { atomicAdd(&g_rd_stats[62], 1ull); atomicAdd(&g_rd_stats[63], (unsigned long long)__popc(act)); } }
#endif
    const int k1 = __ffs(mask) - 1;
    mask &= mask - 1;
    const int k2 = mask ? __ffs(mask) - 1 : k1;
    const bool two = mask != 0;
    mask &= mask - 1;
    PairOps o1, o2;
    pair_ops(k1, o1);
    pair_ops(k2, o2);
    const bool a1ok = ((trivA >> k1) & 1) | in_voronoi(o1.lb, o1.la, o1.nA, o1.tA, o1.AdB, o1.dA, o1.eA);
    const bool b1ok = ((trivB >> k1) & 1) | in_voronoi(o1.la, o1.lb, o1.nB, o1.tB, o1.AdB, o1.dB, o1.eB);
    const bool a2ok = ((trivA >> k2) & 1) | in_voronoi(o2.lb, o2.la, o2.nA, o2.tA, o2.AdB, o2.dA, o2.eA);
    const bool b2ok = ((trivB >> k2) & 1) | in_voronoi(o2.la, o2.lb, o2.nB, o2.tB, o2.AdB, o2.dB, o2.eB);
    const bool ok1 = a1ok && b1ok, ok2 = two && a2ok && b2ok;
    if (ok1 || ok2)
    {
      const PairOps &o = ok1 ? o1 : o2;
      seg_params(t, u, o.la, o.lb, o.AdB, o.dB, o.eB);
      kf = ok1 ? k1 : k2;
#ifdef C2A_RD_STATS
      atomicAdd(&g_rd_stats[ok1 ? 26 : 27], 1ull);
      atomicAdd(&g_rd_stats[32 + kf], 1ull);
#endif
      break;
    }
  }
#ifdef C2A_RD_STATS
  atomicAdd(&g_rd_stats[17 + min(rd_trips, 8)], 1ull);
  if (kf < 0) atomicAdd(&g_rd_stats[28], 1ull);
#endif

  if (kf >= 0)
  {
    // closest points of the accepted pair (the P/Q/S block after each ClosestPoint, :233-848)
    const bool ea1 = kf < 8, eb1 = !((kf >> 2) & 1), ua = !((kf >> 1) & 1), ub = !(kf & 1);
    const double bO = eb1 ? b0 : b1;
    double Q[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      const double cE = eb1 ? R[3 * i + 1] : R[3 * i + 0];
      const double cO = eb1 ? R[3 * i + 0] : R[3 * i + 1];
      Q[i] = ub ? (T[i] + cO * bO + cE * u) : (T[i] + cE * u);
    }
    const double P0 = ea1 ? (ua ? a0 : 0.0) : t;
    const double P1 = ea1 ? t : (ua ? a1 : 0.0);
    S[0] = Q[0] - P0; S[1] = Q[1] - P1; S[2] = Q[2] - 0.0;
    return sqrt(v_dot(S, S));
  }

  // no edge pair: separation along the two face normals (:850-933)
  double sep1, sep2;
  if (T[2] > 0.0)
  {
    sep1 = T[2];
    if (R[6] < 0.0) sep1 += b0 * R[6];
    if (R[7] < 0.0) sep1 += b1 * R[7];
  }
  else
  {
    sep1 = -T[2];
    if (R[6] > 0.0) sep1 -= b0 * R[6];
    if (R[7] > 0.0) sep1 -= b1 * R[7];
  }
  if (Tba[2] < 0)
  {
    sep2 = -Tba[2];
    if (R[2] < 0.0) sep2 += a0 * R[2];
    if (R[5] < 0.0) sep2 += a1 * R[5];
  }
  else
  {
    sep2 = Tba[2];
    if (R[2] > 0.0) sep2 -= a0 * R[2];
    if (R[5] > 0.0) sep2 -= a1 * R[5];
  }
  if (sep1 >= sep2 && sep1 >= 0)
  {
    S[0] = 0.0 - 0.0; S[1] = 0.0 - 0.0;
    S[2] = ((T[2] > 0.0) ? sep1 : -sep1) - 0.0;
  }
  if (sep2 >= sep1 && sep2 >= 0)
  {
    double P[3];
    if (Tba[2] < 0) { P[0] = R[2] * sep2 + T[0]; P[1] = R[5] * sep2 + T[1]; P[2] = R[8] * sep2 + T[2]; }
    else { P[0] = -R[2] * sep2 + T[0]; P[1] = -R[5] * sep2 + T[1]; P[2] = -R[8] * sep2 + T[2]; }
    S[0] = T[0] - P[0]; S[1] = T[1] - P[1]; S[2] = T[2] - P[2];
  }
  const double sep = (sep1 > sep2 ? sep1 : sep2);
  return (sep > 0 ? sep : 0);
}

// ---- triangle-triangle distance ----------------------------------------------------------
// PQP SegPoints (in-tree copy C2A/src/C2A.cpp:59-163), reshaped for SIMT: the reference's nine-way
// branch (u <= 0 | u >= 1 | inside) x (t <= 0 | t >= 1 | inside) becomes selects over values every
// branch spells the same way, so the lanes of a LEAF pass stay converged:
//   Y   = Q | Q + B | Q + B*u            X = P | P + A | P + A*t
//   VEC = Y - X                          both parameters clamped (vertex-vertex)
//       = W x ((TT) x W) reordered as the reference does (TMP = TT x W; VEC = W x TMP), with
//         W = A, TT = Y - P   (u clamped, t inside;   Y = Q gives the reference's T = Q - P)
//         W = B, TT = Q - X   (u inside,  t clamped;  X = P gives the reference's T = Q - P)
//       = +-(A x B)                      both inside
// Each value is computed by the same operations in the same order as in the branch that uses it.
C2A_DEV void seg_points(double VEC[3], double X[3], double Y[3], const double P[3], const double A[3],
                        const double Q[3], const double B[3])
{
  double T[3], TT[3], TMP[3], W[3], Xm[3], Ym[3], Ve[3], Vx[3], Vm[3];
  v_sub(T, Q, P);
  const double AdA = v_dot(A, A), BdB = v_dot(B, B), AdB = v_dot(A, B);
  const double AdT = v_dot(A, T), BdT = v_dot(B, T);
  const double denom = AdA * BdB - AdB * AdB;
  double t = (AdT * BdB - BdT * AdB) / denom;
  t = ((t < 0) || (t != t)) ? 0.0 : ((t > 1) ? 1.0 : t);
  const double u = (t * AdB - BdT) / BdB;
  const bool u0 = (u <= 0) || (u != u), u1 = !u0 && (u >= 1), um = !u0 && !u1;
  const double t2 = (u1 ? (AdB + AdT) : AdT) / AdA;
  t = um ? t : t2;
  const bool t0 = (t <= 0) || (t != t), t1 = !t0 && (t >= 1), tm = !t0 && !t1;

  v_madd(Ym, Q, B, u);
  v_madd(Xm, P, A, t);
#pragma unroll
  for (int k = 0; k < 3; k++)
  {
    Y[k] = u0 ? Q[k] : (u1 ? (Q[k] + B[k]) : Ym[k]);
    X[k] = t0 ? P[k] : (t1 ? (P[k] + A[k]) : Xm[k]);
    W[k] = um ? B[k] : A[k];
  }
  v_sub(Ve, Y, X);
#pragma unroll
  for (int k = 0; k < 3; k++) TT[k] = um ? (Q[k] - X[k]) : (Y[k] - P[k]);
  v_cross(TMP, TT, W);
  v_cross(Vx, W, TMP);
  v_cross(Vm, A, B);
  const bool flip = v_dot(Vm, T) < 0;
#pragma unroll
  for (int k = 0; k < 3; k++)
  {
    const double vm = flip ? -Vm[k] : Vm[k];  // the reference multiplies by -1: the same value
    VEC[k] = (um && tm) ? vm : ((!um && !tm) ? Ve[k] : Vx[k]);
  }
}

// closest pair (face of F, vertex of G) test, C2A/src/C2A.cpp:282-339 and :346-391
C2A_DEV bool face_vertex_case(const double F[9], const double Fv[9], const double G[9], int &shown_disjoint,
                              double onFace[3], double vertex[3])
{
  double n[3], V[3], Z[3], proj[3];
  v_cross(n, &Fv[0], &Fv[3]);
  const double nl = v_dot(n, n);
  if (!(nl > 1e-15)) return false;
  v_sub(V, &F[0], &G[0]); proj[0] = v_dot(V, n);
  v_sub(V, &F[0], &G[3]); proj[1] = v_dot(V, n);
  v_sub(V, &F[0], &G[6]); proj[2] = v_dot(V, n);
  int point = -1;
  if ((proj[0] > 0) && (proj[1] > 0) && (proj[2] > 0))
  {
    point = (proj[0] < proj[1]) ? 0 : 1;
    if (proj[2] < (point ? proj[1] : proj[0])) point = 2;
  }
  else if ((proj[0] < 0) && (proj[1] < 0) && (proj[2] < 0))
  {
    point = (proj[0] > proj[1]) ? 0 : 1;
    if (proj[2] > (point ? proj[1] : proj[0])) point = 2;
  }
  if (point < 0) return false;
  shown_disjoint = 1;
  double g[3];
  const double pp = (point == 0) ? proj[0] : ((point == 1) ? proj[1] : proj[2]);
#pragma unroll
  for (int k = 0; k < 3; k++) g[k] = (point == 0) ? G[k] : ((point == 1) ? G[3 + k] : G[6 + k]);
#pragma unroll
  for (int e = 0; e < 3; e++)
  {
    v_sub(V, g, &F[3 * e]);
    v_cross(Z, n, &Fv[3 * e]);
    if (!(v_dot(V, Z) > 0)) return false;
  }
  v_madd(onFace, g, n, pp / nl);
  v_cpy(vertex, g);
  return true;
}

// rotate three 3-vectors stored back to back: (v0, v1, v2) <- (v1, v2, v0)
C2A_DEV void rot3(double v[9])
{
#pragma unroll
  for (int k = 0; k < 3; k++)
  {
    const double t = v[k];
    v[k] = v[3 + k]; v[3 + k] = v[6 + k]; v[6 + k] = t;
  }
}

// PQP TriDist (in-tree copy C2A/src/C2A.cpp:165-405 without the contact-feature writes).
// The 3x3 edge-pair loop stays rolled (one SegPoints body), but the vertex / edge arrays are never indexed by
// the loop counters: after each trip the vertices are rotated by one, so edge i / j always starts at position 0
// and the opposite vertex (i+2)%3 / (j+2)%3 is at position 2 -- the arrays stay in registers instead of local
// memory (whose lines the small L1 left beside the slot pool keeps losing to L2).  Three rotations restore the
// order.  Edge vectors are re-formed per trip (same subtraction, same bits) rather than kept live.
C2A_DEV double tri_dist(double P[3], double Q[3], const double S_in[9], const double T_in[9])
{
  double S[9], T[9], VEC[3], V[3], Z[3];
#pragma unroll
  for (int k = 0; k < 9; k++) { S[k] = S_in[k]; T[k] = T_in[k]; }

  double minP[3], minQ[3], mindd;
  int shown_disjoint = 0;
  mindd = v_dist2(&S[0], &T[0]) + 1;

#pragma unroll 1
  for (int ij = 0; ij < 9; ij++)
  {
    // edge i of S: S[0..2] + Sv[0..2], opposite vertex S[6..8]; edge j of T likewise
    double A[3], B[3];  // edge vectors Sv[i] = S[i+1] - S[i], Tv[j] = T[j+1] - T[j]: the same subtraction every time
    v_sub(A, &S[3], &S[0]);
    v_sub(B, &T[3], &T[0]);
    seg_points(VEC, P, Q, &S[0], A, &T[0], B);
    v_sub(V, Q, P);
    const double dd = v_dot(V, V);
    if (dd <= mindd)
    {
      v_cpy(minP, P); v_cpy(minQ, Q); mindd = dd;
      v_sub(Z, &S[6], P);
      double a = v_dot(Z, VEC);
      v_sub(Z, &T[6], Q);
      double b = v_dot(Z, VEC);
      if ((a <= 0) && (b >= 0)) return sqrt(dd);
      const double p = v_dot(V, VEC);
      if (a < 0) a = 0;
      if (b > 0) b = 0;
      if ((p - a + b) > 0) shown_disjoint = 1;
    }
    rot3(T);                                     // next j
    if (ij == 2 || ij == 5 || ij == 8) rot3(S);  // j wrapped: next i (T is back in order)
  }

  double Sv[9], Tv[9];
  v_sub(&Sv[0], &S[3], &S[0]); v_sub(&Sv[3], &S[6], &S[3]); v_sub(&Sv[6], &S[0], &S[6]);
  v_sub(&Tv[0], &T[3], &T[0]); v_sub(&Tv[3], &T[6], &T[3]); v_sub(&Tv[6], &T[0], &T[6]);
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

// TriDist with contact features: the reference's in-tree variant (C2A/src/C2A.cpp:165-405) used by the
// contact pass.  f?_type: 0 vertex, 1 edge, 2 face; f?_fid: vertex / edge index (-1 for a face).
C2A_DEV double tri_dist_features(double P[3], double Q[3], const double S[9], const double T[9], int &f1_type,
                                 int &f1_fid, int &f2_type, int &f2_fid)
{
  double Sv[9], Tv[9], VEC[3], V[3], Z[3];
  v_sub(&Sv[0], &S[3], &S[0]); v_sub(&Sv[3], &S[6], &S[3]); v_sub(&Sv[6], &S[0], &S[6]);
  v_sub(&Tv[0], &T[3], &T[0]); v_sub(&Tv[3], &T[6], &T[3]); v_sub(&Tv[6], &T[0], &T[6]);
  double minP[3], minQ[3], mindd;
  int shown_disjoint = 0;
  mindd = v_dist2(&S[0], &T[0]) + 1;
#pragma unroll 1
  for (int i = 0; i < 3; i++)
  {
#pragma unroll 1
    for (int j = 0; j < 3; j++)
    {
      seg_points(VEC, P, Q, &S[3 * i], &Sv[3 * i], &T[3 * j], &Tv[3 * j]);
      v_sub(V, Q, P);
      const double dd = v_dot(V, V);
      if (dd <= mindd)
      {
        v_cpy(minP, P); v_cpy(minQ, Q); mindd = dd;
        f1_type = 1; f1_fid = i; f2_type = 1; f2_fid = j;
        v_sub(Z, &S[3 * ((i + 2) % 3)], P);
        double a = v_dot(Z, VEC);
        v_sub(Z, &T[3 * ((j + 2) % 3)], Q);
        double b = v_dot(Z, VEC);
        if ((a <= 0) && (b >= 0)) return sqrt(dd);
        const double p = v_dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
    }
  }
  // which vertex the vertex-face cases pick (the feature id): the same choice face_vertex_case makes
  auto pick_vertex = [&](const double *F, const double *Fv, const double *Gt) {
    double n[3], W[3], proj[3];
    v_cross(n, &Fv[0], &Fv[3]);
    v_sub(W, &F[0], &Gt[0]); proj[0] = v_dot(W, n);
    v_sub(W, &F[0], &Gt[3]); proj[1] = v_dot(W, n);
    v_sub(W, &F[0], &Gt[6]); proj[2] = v_dot(W, n);
    int point = -1;
    if ((proj[0] > 0) && (proj[1] > 0) && (proj[2] > 0)) { point = (proj[0] < proj[1]) ? 0 : 1; if (proj[2] < (point ? proj[1] : proj[0])) point = 2; }
    else if ((proj[0] < 0) && (proj[1] < 0) && (proj[2] < 0)) { point = (proj[0] > proj[1]) ? 0 : 1; if (proj[2] > (point ? proj[1] : proj[0])) point = 2; }
    return point;
  };
  double onFace[3], vert[3];
  if (face_vertex_case(S, Sv, T, shown_disjoint, onFace, vert))
  {
    f1_type = 2; f1_fid = -1; f2_type = 0; f2_fid = pick_vertex(S, Sv, T);
    v_cpy(P, onFace); v_cpy(Q, vert);
    return sqrt(v_dist2(P, Q));
  }
  if (face_vertex_case(T, Tv, S, shown_disjoint, onFace, vert))
  {
    f1_type = 0; f1_fid = pick_vertex(T, Tv, S); f2_type = 2; f2_fid = -1;
    v_cpy(P, vert); v_cpy(Q, onFace);
    return sqrt(v_dist2(P, Q));
  }
  if (shown_disjoint) { v_cpy(P, minP); v_cpy(Q, minQ); return sqrt(mindd); }
  return 0;
}

// PQP TriDistance: bring triangle 2 into triangle 1's frame, then TriDist.
C2A_DEV double tri_distance(const double R[9], const double T[3], const double t1[9], const double t2[9],
                            double p[3], double q[3])
{
  double tri2[9];
  m_v_p(&tri2[0], R, &t2[0], T);
  m_v_p(&tri2[3], R, &t2[3], T);
  m_v_p(&tri2[6], R, &t2[6], T);
  return tri_dist(p, q, t1, tri2);
}

}  // namespace c2a
