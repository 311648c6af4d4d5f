// Translation-only branch of the CCD query on the device (both angular speeds < 1e-8,
// /root/reference/C2A/src/C2A.cpp:2391-2395):
//   C2A_TimeOfContactStep, b_TanslationCCD paths      C2A/src/C2A.cpp:1818-1852, 1900-1917
//   TOCStepRecurse_Dis_Translation                    C2A/src/C2A.cpp:1362-1521
//   CInterpMotion_Linear::CAonRSS                     C2A/src/InterpMotion.cpp:570-655
//   CInterpMotion_Linear::CAonNonAdjacentTriangles    C2A/src/InterpMotion.cpp:657-743
//   C2A_QueryTimeOfContact, translation exit          C2A/src/C2A.cpp:2032-2049
//
// One traversal per query (no outer CA loop: the advancement runs per node pair, <= 50 inner steps),
// in which only objmotion1's velocity is advanced -- the reference calls objmotion1->CAonRSS /
// CAonNonAdjacentTriangles and never reads objmotion2's cv there: reproduced, not fixed.  The branch is
// rare next to the rotational one (a pure translation of BOTH bodies), so it is a plain
// one-thread-per-query depth-first search with a local stack, launched after the main kernel over the
// same batch; threads whose query has rotation return at once.
//
// Undefined behaviour in the reference and what is done about it: C2ARectDist leaves S untouched when no
// edge pair is accepted and both face separations are negative (crossing rectangles,
// C2A_RectDist.h:887-928) and CAonRSS then reads its uninitialised local S (InterpMotion.cpp:579-592).
// Any finite non-zero garbage gives the same outcome (d = 0, a negative step, "descend iff
// 0 < res->distance"), so S is preset to (1,0,0) -- bit-identical to the reference's object code on every
// fixture.  res->p1/p2 are copied from never-written locals there (C2A.cpp:1392,1411-1412): zero here.
#pragma once
#include "c2a_solve.cuh"

namespace c2a {

struct TransArgs
{
  DevModel A, B;
  const double *motions;  // [n][2][MOTION_DOUBLES]; rec[23] of object 1 = CInterpMotion::m_toc_delta (0: tol_d)
  const int *seedA, *seedB;
  long long n;
  double tol_d;
  c2a_b200_results out;
  const int *order;  // optional [n]: the batch is the queries order[0..n) (heterogeneous batches), NULL = 0..n-1
  double *gstack;    // GS = true: [entries][TRANS_ENTRY][threads] traversal stacks in global memory (hierarchies deeper than TRANS_STACK)
};

constexpr int TRANS_STACK = 96;           // local-memory stack; deeper hierarchies (depth(A)+depth(B)+2 > 96) run the GS = true instance
constexpr int TRANS_ENTRY = 15;           // R(9) T(3) ids mint_child d_child
constexpr double SECURITY_RATIO = 0.1;    // GMP_CCD_SECURITY_DISTANCE_RATIO, InterpMotion.cpp:289

// C2A_BV_Distance, C2A/src/C2A_BV.cpp:666-675
__device__ __noinline__ double bv_distance_nl(const double R[9], const double T[3], const double *ga, const double *gb, double S[3])
{
  double d = rss_rect_dist(R, T, __ldg(ga + 12), __ldg(ga + 13), __ldg(gb + 12), __ldg(gb + 13), S);
  d -= (__ldg(ga + 14) + __ldg(gb + 14));
  return (d < 0.0) ? 0.0 : d;
}

__device__ __noinline__ double tri_dist_nl(double p[3], double q[3], const double S[9], const double T[9])
{
  return tri_dist(p, q, S, T);
}

// CAonRSS (InterpMotion.cpp:570-655).  The first distance test and the ones inside the while loop share
// one call site; bValid_C_clo is always true (C2A_RectDist.h:931) so the centre-of-mass branch is dead.
C2A_DEV bool ca_on_rss(const Motion &m1, double delta, const double r1[9], const double R[9], const double T[3],
                       const double *ga, const double *rl, const double *gb, double *mint, double *distance)
{
  double S[3] = {1.0, 0.0, 0.0}, temp1[3], Tcur[3], Vel[3], Rl[9];
  load9v(Rl, rl);
  mt_v(temp1, r1, m1.cv);
  mt_v(Vel, Rl, temp1);
  const double ang = __ldg(ga + 15);
  v_cpy(Tcur, T);
  double total_toc = 0.0;
  int nIters = 0;
  bool first = true;
  while (true)
  {
    const double d = bv_distance_nl(R, Tcur, ga, gb, S);
    if (!first && d == 0) break;
    m_v(temp1, Rl, S);
    m_v(S, r1, temp1);
    const double tocf = (d - SECURITY_RATIO * delta) / motion_bound_leaf(m1, ang, S);
    if (first) { total_toc = tocf; nIters = 1; first = false; }
    else
    {
      nIters++;
      if (tocf < delta) break;
      total_toc += tocf;
    }
    if (!((d >= delta) && (total_toc <= mint[0]) && (nIters < 50))) break;
    v_madd(Tcur, T, Vel, -total_toc);
  }
  if (total_toc < 1.0)
  {
    mint[0] = total_toc;
    distance[0] = (total_toc - delta) * v_len(m1.cv);
    if (distance[0] < 0) distance[0] = 0;
    return true;
  }
  return false;
}

// CAonNonAdjacentTriangles (InterpMotion.cpp:657-743); triB is already in triangle A's frame
C2A_DEV bool ca_on_triangles(const Motion &m1, double delta, const double r1[9], const double triA[9],
                             const double triB[9], double *mint, double *distance)
{
  double p[3], q[3], At[9], Vel[3], nrm[3];
  mt_v(Vel, r1, m1.cv);
#pragma unroll
  for (int i = 0; i < 9; i++) At[i] = triA[i];
  double total_toc = 0.0;
  int nIters = 0;
  bool first = true;
  while (true)
  {
    const double d = tri_dist_nl(p, q, At, triB);
    if (!first && d == 0.0) break;
    v_sub(nrm, q, p);
    v_normalize(nrm);
    double u = v_dot(Vel, nrm);
    if (u <= 0) u = 1e-30;
    if (first)
    {
      if (d == 0) { mint[0] = 0.0; distance[0] = 0.0; return true; }
      const double dt = (d - SECURITY_RATIO * delta) / u;
      nIters = 1;
      if (dt >= mint[0]) return false;
      total_toc = dt;
      first = false;
    }
    else
    {
      const double tofc = (d - SECURITY_RATIO * delta) / u;
      nIters++;
      if (tofc < delta) break;
      total_toc += tofc;
    }
    if (!((d > delta) && (total_toc <= mint[0]) && (nIters < 50))) break;
#pragma unroll
    for (int i = 0; i < 3; i++) v_madd(&At[3 * i], &triA[3 * i], Vel, total_toc);
  }
  if (total_toc <= mint[0] && total_toc >= 0.0)
  {
    mint[0] = total_toc;
    distance[0] = (total_toc)*v_len(m1.cv) + delta;
    return true;
  }
  return false;
}

// GS: the per-thread stack lives in global memory, entry-major and thread-interleaved (what local memory does, without its
// size limit) -- only used for hierarchies deeper than TRANS_STACK
template <bool GS>
__global__ void __launch_bounds__(128) c2a_translation_kernel(const TransArgs args)
{
  const DevModel &A = args.A, &B = args.B;
  double stk_local[GS ? 1 : TRANS_STACK * TRANS_ENTRY];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#define STK(e, f) (*(GS ? (args.gstack + ((size_t)(e) * TRANS_ENTRY + (f)) * (size_t)stride + gtid) : (stk_local + (e) * TRANS_ENTRY + (f))))
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < args.n; k += stride)
  {
    const long long q = args.order ? (long long)__ldg(args.order + k) : k;
    const double *rec = args.motions + (size_t)(2 * MOTION_DOUBLES) * q;
    if (!(__ldg(rec + 18) < 1e-8 && __ldg(rec + MOTION_DOUBLES + 18) < 1e-8)) continue;  // C2A.cpp:2391-2395
    Motion m1;
    motion_load(m1, rec);
    const double dl = __ldg(rec + 23);
    const double delta = (dl != 0.0) ? dl : args.tol_d;  // m_toc_delta = d_delta, C2A.cpp:2384-2388

    // C2A_TimeOfContactStep at the start poses, :1791-1806
    double r1[9], tt1[3], R2[9], T2[3], Rrel[9], Trel[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3];
    load9(r1, rec); load3(tt1, rec + 9);
    load9(R2, rec + MOTION_DOUBLES); load3(T2, rec + MOTION_DOUBLES + 9);
    mt_m(Rrel, r1, R2);
    v_sub(Tt, T2, tt1);
    mt_v(Trel, r1, Tt);
#pragma unroll
    for (int i = 0; i < 12; i++) { g1[i] = __ldg(A.geom + i); g2[i] = __ldg(B.geom + i); }
    m_m(Rt, Rrel, g2);
    mt_m(R, g1, Rt);
    m_v_p(Tt, Rrel, &g2[9], Trel);
    v_sub(Tt, Tt, &g1[9]);
    mt_v(T, g1, Tt);

    int lastA = seed_or_zero(args.seedA, q, args.A.n_tris), lastB = seed_or_zero(args.seedB, q, args.B.n_tris);  // res->last_triA / last_triB
    double res_mint, res_dist;
    {
      // seed: advancement of the two seed triangles, :1818-1852
      double t1[9], t2[9], tri2[9];
      load9v(t1, A.tris + (size_t)TRI_STRIDE * lastA);
      load9v(t2, B.tris + (size_t)TRI_STRIDE * lastB);
      m_v_p(&tri2[0], Rrel, &t2[0], Trel); m_v_p(&tri2[3], Rrel, &t2[3], Trel); m_v_p(&tri2[6], Rrel, &t2[6], Trel);
      double mint = 1.0, dTri = 0.0;
      if (ca_on_triangles(m1, delta, r1, t1, tri2, &mint, &dTri)) { res_mint = mint; res_dist = dTri; }
      else { res_mint = 1.0; res_dist = 1e+30; }
    }

    int nbv = 0, ntri = 0, sp = 0;
    {
#pragma unroll
      for (int i = 0; i < 9; i++) STK(0, i) = R[i];
      STK(0, 9) = T[0]; STK(0, 10) = T[1]; STK(0, 11) = T[2]; STK(0, 12) = __hiloint2double(0, 0);
      STK(0, 13) = -1.0; STK(0, 14) = -1.0;  // the root pair is visited unconditionally (res->mint, res->distance are >= 0)
      sp = 1;
    }
    while (sp > 0)
    {
      const int ei = sp - 1;
      sp--;
      const double e_mt = STK(ei, 13), e_d = STK(ei, 14);
      // :1481-1515, evaluated with the state at the moment the reference would reach this child
      if (!(e_mt < res_mint && ((e_d < (res_dist - 0.0)) || (e_d * (1 + 0.0) < res_dist)))) continue;
#pragma unroll
      for (int i = 0; i < 9; i++) R[i] = STK(ei, i);
      T[0] = STK(ei, 9); T[1] = STK(ei, 10); T[2] = STK(ei, 11);
      const double e_ids = STK(ei, 12);
      const int b1 = __double2hiint(e_ids), b2 = __double2loint(e_ids);
      const NodeMeta ma = A.meta[b1], mb = B.meta[b2];
      const bool l1 = ma.first_child < 0, l2 = mb.first_child < 0;
      if (l1 && l2)
      {
        // :1390-1420
        const int ta = -ma.first_child - 1, tb = -mb.first_child - 1;
        double t1[9], t2[9], tri2[9];
        load9v(t1, A.tris + (size_t)TRI_STRIDE * ta);
        load9v(t2, B.tris + (size_t)TRI_STRIDE * tb);
        m_v_p(&tri2[0], Rrel, &t2[0], Trel); m_v_p(&tri2[3], Rrel, &t2[3], Trel); m_v_p(&tri2[6], Rrel, &t2[6], Trel);
        if (ca_on_triangles(m1, delta, r1, t1, tri2, &res_mint, &res_dist)) { lastA = ta; lastB = tb; }
        ntri++;
        continue;
      }
      // :1426-1476: both children, each with its own inner advancement
      double mt_ac[2], d_ac[2], ids[2];
      const bool split1 = l2 || (!l1 && (ma.size > mb.size));  // :1432
      double Rch[2][9], Tch[2][3];
#pragma unroll 1
      for (int c = 0; c < 2; c++)
      {
        const double *ga, *gb, *rl;
        if (split1)
        {
          const int n1 = ma.first_child + c;
          ids[c] = __hiloint2double(n1, b2);
          ga = A.geom + (size_t)n1 * GEOM_STRIDE; gb = B.geom + (size_t)b2 * GEOM_STRIDE; rl = A.rloc + (size_t)n1 * RLOC_STRIDE;
          double Rn[9], Tn[3];
          load_node_rt(Rn, Tn, ga);
          mt_m(Rch[c], Rn, R); v_sub(Tt, T, Tn); mt_v(Tch[c], Rn, Tt);
        }
        else
        {
          const int n2 = mb.first_child + c;
          ids[c] = __hiloint2double(b1, n2);
          ga = A.geom + (size_t)b1 * GEOM_STRIDE; gb = B.geom + (size_t)n2 * GEOM_STRIDE; rl = A.rloc + (size_t)b1 * RLOC_STRIDE;
          double Rn[9], Tn[3];
          load_node_rt(Rn, Tn, gb);
          m_m(Rch[c], R, Rn); m_v_p(Tch[c], R, Tn, T);
        }
        mt_ac[c] = res_mint; d_ac[c] = 1e+30;
        ca_on_rss(m1, delta, r1, Rch[c], Tch[c], ga, rl, gb, &mt_ac[c], &d_ac[c]);
      }
      nbv += 2;
      const bool c_first = d_ac[1] < d_ac[0];
      // push the one visited second first
#pragma unroll 1
      for (int k = 0; k < 2; k++)
      {
        const int c = (k == 0) ? (c_first ? 0 : 1) : (c_first ? 1 : 0);
#pragma unroll
        for (int i = 0; i < 9; i++) STK(sp, i) = Rch[c][i];
        STK(sp, 9) = Tch[c][0]; STK(sp, 10) = Tch[c][1]; STK(sp, 11) = Tch[c][2];
        STK(sp, 12) = ids[c];
        STK(sp, 13) = mt_ac[c]; STK(sp, 14) = d_ac[c];
        sp++;
      }
    }

    // :1907-1916: distance of the last improving triangle pair at the pose of the step bound
    double Ra[9], Ta[3], Rb[9], Tb[3], p[3], qq[3];
    motion_pose_nl(rec, res_mint, Ra, Ta);
    motion_pose_nl(rec + MOTION_DOUBLES, res_mint, Rb, Tb);
    mt_m(Rrel, Ra, Rb);
    v_sub(Tt, Tb, Ta);
    mt_v(Trel, Ra, Tt);
    {
      double t1[9], t2[9], tri2[9];
      load9v(t1, A.tris + (size_t)TRI_STRIDE * lastA);
      load9v(t2, B.tris + (size_t)TRI_STRIDE * lastB);
      m_v_p(&tri2[0], Rrel, &t2[0], Trel); m_v_p(&tri2[3], Rrel, &t2[3], Trel); m_v_p(&tri2[6], Rrel, &t2[6], Trel);
      res_dist = tri_dist_nl(p, qq, t1, tri2);
    }

    // C2A_QueryTimeOfContact :2032-2049 and the pose outputs of C2A_Solve :2411-2429
    const c2a_b200_results &o = args.out;
    const bool hit = !(res_mint >= 1.0);
    if (o.status) o.status[q] = C2A_B200_QUERY_OK;
    if (o.collisionfree) o.collisionfree[q] = hit ? 0 : 1;
    if (o.num_ca) o.num_ca[q] = 0;
    if (o.num_bv_tests) o.num_bv_tests[q] = nbv;
    if (o.num_tri_tests) o.num_tri_tests[q] = ntri;
    if (o.toc) o.toc[q] = res_mint;
    if (o.distance) o.distance[q] = res_dist;
    if (o.mint) o.mint[q] = res_mint;
    if (o.last_tri) { o.last_tri[2 * q] = lastA; o.last_tri[2 * q + 1] = lastB; }
    if (o.p1p2)
    {
#pragma unroll
      for (int i = 0; i < 6; i++) o.p1p2[6 * q + i] = 0.0;
    }
    if (hit && o.pose_toc)
    {
      // integrate(toc) after integrate(mint) with toc == mint: the same pose
#pragma unroll
      for (int i = 0; i < 9; i++) { o.pose_toc[24 * q + i] = Ra[i]; o.pose_toc[24 * q + 12 + i] = Rb[i]; }
#pragma unroll
      for (int i = 0; i < 3; i++) { o.pose_toc[24 * q + 9 + i] = Ta[i]; o.pose_toc[24 * q + 21 + i] = Tb[i]; }
    }
  }
#undef STK
}

}  // namespace c2a
