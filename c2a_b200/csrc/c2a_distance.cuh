// Discrete minimum-distance query on the device: C2A_Distance with the depth-first routine it takes for
// qsize <= 2 (/root/reference/C2A/src/C2A_PQP.cpp:970-1056 and C2ADistanceRecurse :481-614).
//
// Not on the CCD hot path; it reuses the hot path's device functions
// (rss_rect_dist, tri_distance).  The result depends on the visiting order exactly as in the CCD traversal
// (res->distance shrinks as leaves are visited and gates the pruning), so a query is one thread walking the
// reference's depth-first order with a local stack; parallelism is across queries.
//
// GATE = true is C2A_Collide's C2A_DistanceResult overload (C2A_PQP.cpp:1060-1280): the same walk, except that a node
// pair whose boxes do not overlap is left at once (C2A_BV_Overlap at the top of C2ACollideRecurse, :1068).  The gate is
// handed the transform this walk chains through the RSS corners Tr, as the reference does (its RSS_TYPE branch comes
// first, :1118-1152), with the boxes' half-dimensions BV::d.
#pragma once
#include "c2a_collide.cuh"

namespace c2a {

struct DistanceArgs
{
  DevModel A, B;
  const double *poses;        // [n][24] pose of A, pose of B (R(9)+T(3) each)
  const int *seedA, *seedB;   // [n] or NULL (triangle 0): o1->last_tri / o2->last_tri going in
  long long n;
  double rel_err, abs_err;
  double *distance;           // [n]
  double *p1p2;               // [n][6] or NULL: closest points, each in its own model's frame
  int *tri_pair;              // [n][2] or NULL: closest triangle pair, builder order (o->last_tri coming out)
  int *num_bv_tests, *num_tri_tests;  // [n] or NULL
  double *gstack;             // GS = true: [entries][DIST_ENTRY][threads] traversal stacks in global memory
  const double *obbA, *obbB;  // GATE = true: [n_nodes][OBB_STRIDE] box half-dimensions d(3) (+ centre, unused here)
};

constexpr int DIST_STACK = 96;  // local-memory stack; deeper hierarchies run the GS = true instance
constexpr int DIST_ENTRY = 14;  // R(9) T(3) ids d

template <bool GS, bool GATE>
__global__ void __launch_bounds__(128) c2a_distance_kernel(const DistanceArgs args)
{
  const DevModel &A = args.A, &B = args.B;
  double stk_local[GS ? 1 : DIST_STACK * DIST_ENTRY];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#define STK(e, f) (*(GS ? (args.gstack + ((size_t)(e) * DIST_ENTRY + (f)) * (size_t)stride + gtid) : (stk_local + (e) * DIST_ENTRY + (f))))
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < args.n; q += stride)
  {
    const double *pose = args.poses + 24 * q;
    double R1[9], T1[3], R2[9], T2[3], Rrel[9], Trel[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3];
    load9(R1, pose); load3(T1, pose + 9); load9(R2, pose + 12); load3(T2, pose + 21);
    // [R,T] = [R1'R2, R1'(T2-T1)], :987-990
    mt_m(Rrel, R1, R2);
    v_sub(Tt, T2, T1);
    mt_v(Trel, R1, Tt);
    int ta = seed_or_zero(args.seedA, q, args.A.n_tris), tb = seed_or_zero(args.seedB, q, args.B.n_tris);
    double p1[3], p2[3];
    // initial upper bound from the last closest triangle pair, :995-1000
    double dist = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * ta, B.tris + (size_t)TRI_STRIDE * tb, p1, p2);
    // root pair, :1016-1027
#pragma unroll
    for (int i = 0; i < 12; i++) { g1[i] = __ldg(A.geom + i); g2[i] = __ldg(B.geom + i); }
    m_m(Rt, Rrel, g2);
    mt_m(R, g1, Rt);
    m_v_p(Tt, Rrel, &g2[9], Trel);
    v_sub(Tt, Tt, &g1[9]);
    mt_v(T, g1, Tt);
    int nbv = 0, ntri = 0, sp = 0;
    {
#pragma unroll
      for (int i = 0; i < 9; i++) STK(0, i) = R[i];
      STK(0, 9) = T[0]; STK(0, 10) = T[1]; STK(0, 11) = T[2]; STK(0, 12) = __hiloint2double(0, 0);
      STK(0, 13) = -1.0;  // flag: the root pair is visited unconditionally
      sp = 1;
    }
    while (sp > 0)
    {
      const int ei = sp - 1;
      sp--;
      const double ed = STK(ei, 13);
      // the descend test, evaluated when the reference would reach this child (:589-613)
      if (ed >= 0.0 && !((ed < (dist - args.abs_err)) || (ed * (1 + args.rel_err) < dist))) continue;
#pragma unroll
      for (int i = 0; i < 9; i++) R[i] = STK(ei, i);
      T[0] = STK(ei, 9); T[1] = STK(ei, 10); T[2] = STK(ei, 11);
      const double e_ids = STK(ei, 12);
      const int b1 = __double2hiint(e_ids), b2 = __double2loint(e_ids);
      if (GATE)
      {
        double da[3], db[3];
#pragma unroll
        for (int i = 0; i < 3; i++) { da[i] = __ldg(args.obbA + (size_t)b1 * OBB_STRIDE + i); db[i] = __ldg(args.obbB + (size_t)b2 * OBB_STRIDE + i); }
        if (obb_disjoint(R, T, da, db) != 0) continue;
      }
      const NodeMeta ma = A.meta[b1], mb = B.meta[b2];
      const bool l1 = ma.first_child < 0, l2 = mb.first_child < 0;
      if (l1 && l2)
      {
        // :494-521
        ntri++;
        const int t1 = -ma.first_child - 1, t2 = -mb.first_child - 1;
        double p[3], qq[3];
        const double d = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * t1, B.tris + (size_t)TRI_STRIDE * t2, p, qq);
        if (d < dist)
        {
          dist = d; ta = t1; tb = t2;
          v_cpy(p1, p); v_cpy(p2, qq);
        }
        continue;
      }
      // :527-587: both children, the nearer one first
      double Rch[2][9], Tch[2][3], dch[2], ids[2];
      const bool split1 = l2 || (!l1 && (ma.size > mb.size));
#pragma unroll 1
      for (int c = 0; c < 2; c++)
      {
        const double *ga, *gb;
        if (split1)
        {
          const int n1 = ma.first_child + c;
          ids[c] = __hiloint2double(n1, b2);
          ga = A.geom + (size_t)n1 * GEOM_STRIDE; gb = B.geom + (size_t)b2 * GEOM_STRIDE;
          double Rn[9], Tn[3];
          load_node_rt(Rn, Tn, ga);
          mt_m(Rch[c], Rn, R); v_sub(Tt, T, Tn); mt_v(Tch[c], Rn, Tt);
        }
        else
        {
          const int n2 = mb.first_child + c;
          ids[c] = __hiloint2double(b1, n2);
          ga = A.geom + (size_t)b1 * GEOM_STRIDE; gb = B.geom + (size_t)n2 * GEOM_STRIDE;
          double Rn[9], Tn[3];
          load_node_rt(Rn, Tn, gb);
          m_m(Rch[c], R, Rn); m_v_p(Tch[c], R, Tn, T);
        }
        double S[3];
        double d = rss_rect_dist(Rch[c], Tch[c], __ldg(ga + 12), __ldg(ga + 13), __ldg(gb + 12), __ldg(gb + 13), S);
        d -= (__ldg(ga + 14) + __ldg(gb + 14));
        dch[c] = (d < 0.0) ? 0.0 : d;
      }
      nbv += 2;
      const bool c_first = dch[1] < dch[0];
#pragma unroll 1
      for (int k = 0; k < 2; k++)
      {
        const int c = (k == 0) ? (c_first ? 0 : 1) : (c_first ? 1 : 0);  // the one visited second is pushed first
#pragma unroll
        for (int i = 0; i < 9; i++) STK(sp, i) = Rch[c][i];
        STK(sp, 9) = Tch[c][0]; STK(sp, 10) = Tch[c][1]; STK(sp, 11) = Tch[c][2];
        STK(sp, 12) = ids[c]; STK(sp, 13) = dch[c];
        sp++;
      }
    }
    args.distance[q] = dist;
    if (args.p1p2)
    {
      // res->p2 is in cs 1; transform it to cs 2, :1044-1048
      double u[3], p2b[3];
      v_sub(u, p2, Trel);
      mt_v(p2b, Rrel, u);
#pragma unroll
      for (int i = 0; i < 3; i++) { args.p1p2[6 * q + i] = p1[i]; args.p1p2[6 * q + 3 + i] = p2b[i]; }
    }
    if (args.tri_pair) { args.tri_pair[2 * q] = ta; args.tri_pair[2 * q + 1] = tb; }
    if (args.num_bv_tests) args.num_bv_tests[q] = nbv;
    if (args.num_tri_tests) args.num_tri_tests[q] = ntri;
  }
#undef STK
}

// C2A_Distance with qsize > 2: C2ADistanceQueueRecurse (C2A_PQP.cpp:624-787) -- best-first over a bounded queue of
// pending node pairs; when the queue cannot take two more (qsize - 1 pending) the routine calls itself on the current pair
// with a fresh queue and carries on with the old one afterwards.  One thread per query; the recursion is a stack of queue
// frames in global memory (a frame per call, at most depth(A) + depth(B) + 2 of them: every call's root is deeper than its
// caller's).  The queue is PQP's BVTQ (not in the reference's tree); the one property that shows in the results is which
// of several equally distant pending pairs leaves first: the earliest queued, as in the stand-in the compiled reference
// is linked with (oracle/pqp_shim/BVTQ.h) -- entries carry a sequence number, so removal need not keep them in order.
struct DistanceQueueArgs
{
  DistanceArgs d;
  int qsize;          // > 2
  int frames;         // depth(A) + depth(B) + 2
  double *arena;      // [threads][frames][1 + qsize * DQ_ENTRY]: per frame the number of pending pairs, then the pairs
};
constexpr int DQ_ENTRY = 15;  // R(9) T(3) ids d seq

__global__ void __launch_bounds__(128) c2a_distance_queue_kernel(const DistanceQueueArgs qa)
{
  const DistanceArgs &args = qa.d;
  const DevModel &A = args.A, &B = args.B;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t frame_doubles = 1 + (size_t)qa.qsize * DQ_ENTRY;
  double *const mine = qa.arena + gtid * (size_t)qa.frames * frame_doubles;
#define CNT(f) (mine[(size_t)(f) * frame_doubles])
#define QE(f, i, k) (mine[(size_t)(f) * frame_doubles + 1 + (size_t)(i) * DQ_ENTRY + (k)])
  for (long long q = (long long)gtid; q < args.n; q += stride)
  {
    const double *pose = args.poses + 24 * q;
    double R1[9], T1[3], R2[9], T2[3], Rrel[9], Trel[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3];
    load9(R1, pose); load3(T1, pose + 9); load9(R2, pose + 12); load3(T2, pose + 21);
    mt_m(Rrel, R1, R2);
    v_sub(Tt, T2, T1);
    mt_v(Trel, R1, Tt);
    int ta = seed_or_zero(args.seedA, q, args.A.n_tris), tb = seed_or_zero(args.seedB, q, args.B.n_tris);
    double p1[3], p2[3];
    double dist = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * ta, B.tris + (size_t)TRI_STRIDE * tb, p1, p2);
#pragma unroll
    for (int i = 0; i < 12; i++) { g1[i] = __ldg(A.geom + i); g2[i] = __ldg(B.geom + i); }
    m_m(Rt, Rrel, g2);
    mt_m(R, g1, Rt);
    m_v_p(Tt, Rrel, &g2[9], Trel);
    v_sub(Tt, Tt, &g1[9]);
    mt_v(T, g1, Tt);
    int nbv = 0, ntri = 0, f = 0, b1 = 0, b2 = 0;
    double seq = 0.0;
    CNT(0) = 0.0;
    while (true)
    {
      // ---- the pair in hand: (b1, b2) placed by (R, T)
      const NodeMeta ma = A.meta[b1], mb = B.meta[b2];
      const bool l1 = ma.first_child < 0, l2 = mb.first_child < 0;
      if (l1 && l2)
      {
        ntri++;
        const int t1 = -ma.first_child - 1, t2 = -mb.first_child - 1;
        double p[3], qq[3];
        const double d = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * t1, B.tris + (size_t)TRI_STRIDE * t2, p, qq);
        if (d < dist)
        {
          dist = d; ta = t1; tb = t2;
          v_cpy(p1, p); v_cpy(p2, qq);
        }
      }
      else if ((int)CNT(f) == qa.qsize - 1)
      {
        // the queue cannot take two more: the routine calls itself on this pair with a fresh queue (:675-680)
        f++;
        CNT(f) = 0.0;
        continue;
      }
      else
      {
        const bool split1 = l2 || (!l1 && (ma.size > mb.size));
        nbv += 2;
#pragma unroll 1
        for (int c = 0; c < 2; c++)
        {
          const double *ga, *gb;
          double Rc[9], Tc[3], ids;
          if (split1)
          {
            const int n1 = ma.first_child + c;
            ids = __hiloint2double(n1, b2);
            ga = A.geom + (size_t)n1 * GEOM_STRIDE; gb = B.geom + (size_t)b2 * GEOM_STRIDE;
            double Rn[9], Tn[3];
            load_node_rt(Rn, Tn, ga);
            mt_m(Rc, Rn, R); v_sub(Tt, T, Tn); mt_v(Tc, Rn, Tt);
          }
          else
          {
            const int n2 = mb.first_child + c;
            ids = __hiloint2double(b1, n2);
            ga = A.geom + (size_t)b1 * GEOM_STRIDE; gb = B.geom + (size_t)n2 * GEOM_STRIDE;
            double Rn[9], Tn[3];
            load_node_rt(Rn, Tn, gb);
            m_m(Rc, R, Rn); m_v_p(Tc, R, Tn, T);
          }
          double S[3];
          double d = rss_rect_dist(Rc, Tc, __ldg(ga + 12), __ldg(ga + 13), __ldg(gb + 12), __ldg(gb + 13), S);
          d -= (__ldg(ga + 14) + __ldg(gb + 14));
          d = (d < 0.0) ? 0.0 : d;
          const int at = (int)CNT(f);
#pragma unroll
          for (int i = 0; i < 9; i++) QE(f, at, i) = Rc[i];
          QE(f, at, 9) = Tc[0]; QE(f, at, 10) = Tc[1]; QE(f, at, 11) = Tc[2];
          QE(f, at, 12) = ids; QE(f, at, 13) = d; QE(f, at, 14) = seq;
          seq += 1.0;
          CNT(f) = (double)(at + 1);
        }
      }
      // ---- the next pair: the nearest pending one of this call; a call whose queue is empty -- or whose nearest pair
      // cannot improve the distance any more (:776-780) -- returns to its caller's queue
      bool done = false;
      while (true)
      {
        const int cnt = (int)CNT(f);
        if (cnt == 0)
        {
          if (f == 0) { done = true; break; }
          f--;
          continue;
        }
        int k = 0;
        double dk = QE(f, 0, 13), sk = QE(f, 0, 14);
        for (int i = 1; i < cnt; i++)
        {
          const double di = QE(f, i, 13), si = QE(f, i, 14);
          if (di < dk || (di == dk && si < sk)) { k = i; dk = di; sk = si; }
        }
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = QE(f, k, i);
        T[0] = QE(f, k, 9); T[1] = QE(f, k, 10); T[2] = QE(f, k, 11);
        const double e_ids = QE(f, k, 12);
        b1 = __double2hiint(e_ids); b2 = __double2loint(e_ids);
        if (k != cnt - 1)
#pragma unroll
          for (int i = 0; i < DQ_ENTRY; i++) QE(f, k, i) = QE(f, cnt - 1, i);
        CNT(f) = (double)(cnt - 1);
        if ((dk + args.abs_err >= dist) && ((dk * (1 + args.rel_err)) >= dist)) { CNT(f) = 0.0; continue; }
        break;
      }
      if (done) break;
    }
    args.distance[q] = dist;
    if (args.p1p2)
    {
      double u[3], p2b[3];
      v_sub(u, p2, Trel);
      mt_v(p2b, Rrel, u);
#pragma unroll
      for (int i = 0; i < 3; i++) { args.p1p2[6 * q + i] = p1[i]; args.p1p2[6 * q + 3 + i] = p2b[i]; }
    }
    if (args.tri_pair) { args.tri_pair[2 * q] = ta; args.tri_pair[2 * q + 1] = tb; }
    if (args.num_bv_tests) args.num_bv_tests[q] = nbv;
    if (args.num_tri_tests) args.num_tri_tests[q] = ntri;
  }
#undef CNT
#undef QE
}

}  // namespace c2a
