// Discrete collision queries on the device: C2A_Collide, both overloads
//   C2A_Collide(PQP_CollideResult*, ..., flag)                 /root/reference/C2A/src/C2A_PQP.cpp:798-968
//   C2A_Collide(C2A_DistanceResult*, ..., rel_err, abs_err)    /root/reference/C2A/src/C2A_PQP.cpp:1060-1280
//     (the second one is the distance walk of c2a_distance.cuh behind this file's box-overlap gate)
//
// Not on the CCD hot path.  A query is one thread walking the reference's depth-first
// order with a stack; parallelism is across queries.
//
// The box-overlap and triangle-overlap tests are PQP's (obb_disjoint, TriContact), which the reference links from
// outside its tree: they are restated from the published separating-axis formulations in the arithmetic of the test
// shim the compiled reference is linked against (oracle/pqp_shim/pqp_shim.cpp, oracle/c2a_oracle.cpp
// orc_obb_disjoint / orc_tri_contact), so parity for these two functions is pinned to that shim, not to PQP itself.
#pragma once
#include "c2a_solve.cuh"

namespace c2a {

// 15-axis separating-axis test of two boxes: half-dimensions a, b; (B, T) places box 2 in box 1's frame.
// Returns the number of the first separating axis, 0 when the boxes overlap.
C2A_DEV int obb_disjoint(const double B[9], const double T[3], const double a[3], const double b[3])
{
  const double reps = 1e-6;
  double Bf[9];
#pragma unroll
  for (int i = 0; i < 9; i++) Bf[i] = fabs(B[i]) + reps;
#pragma unroll
  for (int i = 0; i < 3; i++)
  {
    const double t = fabs(T[i]);
    if (t > a[i] + b[0] * Bf[3 * i + 0] + b[1] * Bf[3 * i + 1] + b[2] * Bf[3 * i + 2]) return 1 + i;
  }
#pragma unroll
  for (int j = 0; j < 3; j++)
  {
    const double s = T[0] * B[0 + j] + T[1] * B[3 + j] + T[2] * B[6 + j];
    const double t = fabs(s);
    if (t > b[j] + a[0] * Bf[0 + j] + a[1] * Bf[3 + j] + a[2] * Bf[6 + j]) return 4 + j;
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double s = T[i2] * B[3 * i1 + j] - T[i1] * B[3 * i2 + j];
      const double t = fabs(s);
      const double ra = a[i1] * Bf[3 * i2 + j] + a[i2] * Bf[3 * i1 + j];
      const double rb = b[j1] * Bf[3 * i + j2] + b[j2] * Bf[3 * i + j1];
      if (t > ra + rb) return 7 + 3 * i + j;
    }
  return 0;
}

// true when the projections of the two vertex triples on ax do not overlap
C2A_DEV bool axis_separates(const double ax[3], const double p[3][3], const double q[3][3])
{
  double pmin = v_dot(ax, p[0]), pmax = pmin, qmin = v_dot(ax, q[0]), qmax = qmin;
#pragma unroll
  for (int i = 1; i < 3; i++)
  {
    const double v = v_dot(ax, p[i]); if (v < pmin) pmin = v; if (v > pmax) pmax = v;
    const double w = v_dot(ax, q[i]); if (w < qmin) qmin = w; if (w > qmax) qmax = w;
  }
  return (pmin > qmax) || (qmin > pmax);
}

// 17-axis separating-axis test of two triangles given in one frame (P: p1,p2,p3 of triangle 1; Q likewise)
C2A_DEV bool tri_contact(const double P[9], const double Q[9])
{
  double p[3][3], q[3][3], e[3][3], f[3][3], n[3], m[3], ax[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { v_sub(p[k], &P[3 * k], &P[0]); v_sub(q[k], &Q[3 * k], &P[0]); }
  v_sub(e[0], p[1], p[0]); v_sub(e[1], p[2], p[1]); v_sub(e[2], p[0], p[2]);
  v_sub(f[0], q[1], q[0]); v_sub(f[1], q[2], q[1]); v_sub(f[2], q[0], q[2]);
  v_cross(n, e[0], e[1]);
  v_cross(m, f[0], f[1]);
  if (axis_separates(n, p, q)) return false;
  if (axis_separates(m, p, q)) return false;
#pragma unroll 1
  for (int i = 0; i < 3; i++)
#pragma unroll 1
    for (int j = 0; j < 3; j++)
    {
      v_cross(ax, e[i], f[j]);
      if (axis_separates(ax, p, q)) return false;
    }
#pragma unroll 1
  for (int i = 0; i < 3; i++)
  {
    v_cross(ax, e[i], n); if (axis_separates(ax, p, q)) return false;
    v_cross(ax, f[i], m); if (axis_separates(ax, p, q)) return false;
  }
  return true;
}

struct CollideArgs
{
  DevModel A, B;
  const double *obbA, *obbB;  // [n_nodes][OBB_STRIDE]: d(3), To(3)
  const double *poses;        // [n][24] pose of A, pose of B (R(9)+T(3) each)
  long long n;
  int flag;                   // 1: all contacts, 2: first contact (C2A/C2A.h:246-247)
  int max_pairs;
  int *num_pairs;             // [n] pairs found (may exceed max_pairs: the rest is counted, not stored)
  int *pairs;                 // [n][max_pairs][2] or NULL: builder-order triangle indices, traversal order
  int *num_bv_tests, *num_tri_tests;  // [n] or NULL
  double *gstack;             // GS = true: [entries][COLL_ENTRY][threads] traversal stacks in global memory
};

constexpr int COLL_STACK = 96;  // local-memory stack; deeper hierarchies run the GS = true instance
constexpr int COLL_ENTRY = 13;  // R(9) T(3) ids

template <bool GS>
__global__ void __launch_bounds__(128) c2a_collide_kernel(const CollideArgs args)
{
  const DevModel &A = args.A, &B = args.B;
  double stk_local[GS ? 1 : COLL_STACK * COLL_ENTRY];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#define STK(e, f) (*(GS ? (args.gstack + ((size_t)(e) * COLL_ENTRY + (f)) * (size_t)stride + gtid) : (stk_local + (e) * COLL_ENTRY + (f))))
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < args.n; q += stride)
  {
    const double *pose = args.poses + 24 * q;
    double R1[9], T1[3], R2[9], T2[3], Rrel[9], Trel[3], Tt[3], Rt[9], R[9], T[3];
    load9(R1, pose); load3(T1, pose + 9); load9(R2, pose + 12); load3(T2, pose + 21);
    // [R,T] = [R1'R2, R1'(T2-T1)], :937-940
    mt_m(Rrel, R1, R2);
    v_sub(Tt, T2, T1);
    mt_v(Trel, R1, Tt);
    // root pair, chained through the box centres To, :946-957
    {
      double g1[9], g2[9], to1[3], to2[3];
#pragma unroll
      for (int i = 0; i < 9; i++) { g1[i] = __ldg(A.geom + i); g2[i] = __ldg(B.geom + i); }
#pragma unroll
      for (int i = 0; i < 3; i++) { to1[i] = __ldg(args.obbA + 3 + i); to2[i] = __ldg(args.obbB + 3 + i); }
      m_m(Rt, Rrel, g2);
      mt_m(R, g1, Rt);
      m_v_p(Tt, Rrel, to2, Trel);
      v_sub(Tt, Tt, to1);
      mt_v(T, g1, Tt);
    }
    int nbv = 0, ntri = 0, npairs = 0, sp = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) STK(0, i) = R[i];
    STK(0, 9) = T[0]; STK(0, 10) = T[1]; STK(0, 11) = T[2]; STK(0, 12) = __hiloint2double(0, 0);
    sp = 1;
    while (sp > 0)
    {
      // CollideRecurse, :798-909
      sp--;
#pragma unroll
      for (int i = 0; i < 9; i++) R[i] = STK(sp, i);
      T[0] = STK(sp, 9); T[1] = STK(sp, 10); T[2] = STK(sp, 11);
      const double e_ids = STK(sp, 12);
      const int b1 = __double2hiint(e_ids), b2 = __double2loint(e_ids);
      nbv++;
      double da[3], db[3];
#pragma unroll
      for (int i = 0; i < 3; i++) { da[i] = __ldg(args.obbA + (size_t)b1 * OBB_STRIDE + i); db[i] = __ldg(args.obbB + (size_t)b2 * OBB_STRIDE + i); }
      if (obb_disjoint(R, T, da, db) != 0) continue;
      const NodeMeta ma = A.meta[b1], mb = B.meta[b2];
      const bool l1 = ma.first_child < 0, l2 = mb.first_child < 0;
      if (l1 && l2)
      {
        ntri++;
        const int t1 = -ma.first_child - 1, t2 = -mb.first_child - 1;
        double P[9], Q[9], v[3];
        const double *pa = A.tris + (size_t)TRI_STRIDE * t1, *pb = B.tris + (size_t)TRI_STRIDE * t2;
#pragma unroll
        for (int i = 0; i < 9; i++) P[i] = __ldg(pa + i);
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
          v[0] = __ldg(pb + 3 * k); v[1] = __ldg(pb + 3 * k + 1); v[2] = __ldg(pb + 3 * k + 2);
          m_v_p(&Q[3 * k], Rrel, v, Trel);
        }
        if (tri_contact(P, Q))
        {
          if (args.pairs && npairs < args.max_pairs)
          {
            int *o = args.pairs + ((size_t)q * args.max_pairs + npairs) * 2;
            o[0] = t1; o[1] = t2;
          }
          npairs++;
          if (args.flag == 2) break;  // PQP_FIRST_CONTACT: every level returns once a pair is known (:877, :901)
        }
        continue;
      }
      const bool split1 = l2 || (!l1 && (ma.size > mb.size));
      // the second child is pushed first: it is visited after the first one's whole subtree
#pragma unroll 1
      for (int c = 1; c >= 0; c--)
      {
        double Rn[9], Tn[3], Rc[9], Tc[3], to[3];
        if (split1)
        {
          const int n1 = ma.first_child + c;
          load_node_rt(Rn, Tn, A.geom + (size_t)n1 * GEOM_STRIDE);
#pragma unroll
          for (int i = 0; i < 3; i++) to[i] = __ldg(args.obbA + (size_t)n1 * OBB_STRIDE + 3 + i);
          mt_m(Rc, Rn, R); v_sub(Tt, T, to); mt_v(Tc, Rn, Tt);
          STK(sp, 12) = __hiloint2double(n1, b2);
        }
        else
        {
          const int n2 = mb.first_child + c;
          load_node_rt(Rn, Tn, B.geom + (size_t)n2 * GEOM_STRIDE);
#pragma unroll
          for (int i = 0; i < 3; i++) to[i] = __ldg(args.obbB + (size_t)n2 * OBB_STRIDE + 3 + i);
          m_m(Rc, R, Rn); m_v_p(Tc, R, to, T);
          STK(sp, 12) = __hiloint2double(b1, n2);
        }
#pragma unroll
        for (int i = 0; i < 9; i++) STK(sp, i) = Rc[i];
        STK(sp, 9) = Tc[0]; STK(sp, 10) = Tc[1]; STK(sp, 11) = Tc[2];
        sp++;
      }
    }
    args.num_pairs[q] = npairs;
    if (args.num_bv_tests) args.num_bv_tests[q] = nbv;
    if (args.num_tri_tests) args.num_tri_tests[q] = ntri;
  }
#undef STK
}

}  // namespace c2a
