// Contact pass of C2A_Solve on the device: C2A_QueryContact -> C2A_TimeOfContactStep_Contact ->
// TOCStepRecurse_Dis_contact (/root/reference/C2A/src/C2A.cpp:1937-1966, 1727-1775, 1523-1725).
//
// A fixed-threshold traversal at the time-of-contact pose: every BV pair closer than the threshold is
// descended (nearer child first, as the reference), every triangle pair within it is reported with its
// contact features.  The threshold does not change during the pass, so there is no order dependence
// in WHAT is found; the visiting order is kept so that the output order is the reference's (reversed:
// it push_front()s into a std::list).  The pass is small next to the TOC search (the threshold is
// 2*distance + 0.001), so it is a plain one-thread-per-query depth-first search with a local stack.
#pragma once
#include "c2a_solve.cuh"

namespace c2a {

struct ContactArgs
{
  DevModel A, B;
  const int *vidxA, *vidxB;     // [n_tris][3] or NULL
  const double *poses;          // [n][24]
  const double *threshold;      // [n] explicit thresholds, or NULL: 2*distance[i] + 0.001 (C2A.cpp:2433)
  const double *distance;       // [n] (used when threshold == NULL)
  const int *collisionfree;     // [n] or NULL: queries with collisionfree != 0 get no pass
  const int *status;            // [n] or NULL: queries with status != 0 get no pass
  long long n;
  int max_contacts;
  int *num_contact;             // [n]
  c2a_b200_contact *contacts;   // [n][max_contacts] or NULL
  unsigned long long *counter;
  double *gstack;               // GS = true: [entries][13][threads] traversal stacks in global memory
};

constexpr int CONTACT_STACK = 96;  // local-memory stack; deeper hierarchies run the GS = true instance
constexpr int CONTACT_ENTRY = 13;  // R(9) T(3) ids

C2A_DEV void feature_ids(int out[3], int type, int fid, const int *v)
{
  out[0] = out[1] = out[2] = -1;
  if (!v) return;
  if (type == 0) out[0] = v[fid];
  else if (type == 1) { out[0] = v[fid]; out[1] = v[(fid + 1) % 3]; }
  else if (type == 2) { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; }
}

template <bool GS>
__global__ void __launch_bounds__(128) c2a_contact_kernel(const ContactArgs args)
{
  const DevModel &A = args.A, &B = args.B;
  double stk_local[GS ? 1 : CONTACT_STACK * CONTACT_ENTRY];
  const size_t gstride = (size_t)gridDim.x * blockDim.x, gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#define STK(e, f) (*(GS ? (args.gstack + ((size_t)(e) * CONTACT_ENTRY + (f)) * gstride + gtid) : (stk_local + (e) * CONTACT_ENTRY + (f))))
  while (true)
  {
    const long long q = (long long)atomicAdd(args.counter, 1ull);
    if (q >= args.n) break;
    int count = 0;
    const bool skip = (args.collisionfree && args.collisionfree[q] != 0) || (args.status && args.status[q] != 0);
    if (!skip)
    {
      const double thr = args.threshold ? args.threshold[q] : 2 * args.distance[q] + 0.001;
      const double *pose = args.poses + 24 * q;
      double R1[9], T1[3], R2[9], T2[3], Rrel[9], Trel[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3];
      load9(R1, pose); load3(T1, pose + 9); load9(R2, pose + 12); load3(T2, pose + 21);
      // C2A_TimeOfContactStep_Contact, :1741-1752
      mt_m(Rrel, R1, R2);
      v_sub(Tt, T2, T1);
      mt_v(Trel, R1, Tt);
#pragma unroll
      for (int i = 0; i < 12; i++) { g1[i] = __ldg(A.geom + i); g2[i] = __ldg(B.geom + i); }
      m_m(Rt, Rrel, g2);
      mt_m(R, g1, Rt);
      m_v_p(Tt, Rrel, &g2[9], Trel);
      v_sub(Tt, Tt, &g1[9]);
      mt_v(T, g1, Tt);
      int sp = 0;
      {
#pragma unroll
        for (int i = 0; i < 9; i++) STK(0, i) = R[i];
        STK(0, 9) = T[0]; STK(0, 10) = T[1]; STK(0, 11) = T[2]; STK(0, 12) = __hiloint2double(0, 0);
        sp = 1;
      }
      while (sp > 0)
      {
        const int ei = sp - 1;
        sp--;
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = STK(ei, i);
        T[0] = STK(ei, 9); T[1] = STK(ei, 10); T[2] = STK(ei, 11);
        const double e_ids = STK(ei, 12);
        const int b1 = __double2hiint(e_ids), b2 = __double2loint(e_ids);
        const NodeMeta ma = A.meta[b1], mb = B.meta[b2];
        const bool l1 = ma.first_child < 0, l2 = mb.first_child < 0;
        if (l1 && l2)
        {
          // :1544-1636
          const int ta = -ma.first_child - 1, tb = -mb.first_child - 1;
          double t1[9], t2[9], tri2[9], p[3], qq[3];
          load9v(t1, A.tris + (size_t)TRI_STRIDE * ta);
          load9v(t2, B.tris + (size_t)TRI_STRIDE * tb);
          m_v_p(&tri2[0], Rrel, &t2[0], Trel); m_v_p(&tri2[3], Rrel, &t2[3], Trel); m_v_p(&tri2[6], Rrel, &t2[6], Trel);
          int f1t = -1, f1f = 0, f2t = -1, f2f = 0;
          const double d = tri_dist_features(p, qq, t1, tri2, f1t, f1f, f2t, f2f);
          if (f1t != -1 && f2t != -1 && d <= thr)
          {
            if (args.contacts && count < args.max_contacts)
            {
              c2a_b200_contact c;
              c.type_a = f1t + 1; c.type_b = f2t + 1;
              feature_ids(c.fid_a, f1t, f1f, args.vidxA ? args.vidxA + 3 * (size_t)ta : nullptr);
              feature_ids(c.fid_b, f2t, f2f, args.vidxB ? args.vidxB + 3 * (size_t)tb : nullptr);
              c.tri_a = ta; c.tri_b = tb;
              v_cpy(c.pa, p);
              double tmp[3];
              v_sub(tmp, qq, Trel);
              mt_v(c.pb, Rrel, tmp);
              c.dist = d;
              args.contacts[(size_t)q * args.max_contacts + count] = c;
            }
            count++;
          }
          continue;
        }
        // :1646-1722: both children, nearer first
        double Ra[9], Ta[3], Rc[9], Tc[3], S[3];
        int a1, a2, c1, c2;
        const double *ga, *gb_a, *gc, *gb_c;
        if (l2 || (!l1 && (ma.size > mb.size)))
        {
          a1 = ma.first_child; a2 = b2; c1 = a1 + 1; c2 = b2;
          ga = A.geom + (size_t)a1 * GEOM_STRIDE; gc = ga + GEOM_STRIDE; gb_a = gb_c = B.geom + (size_t)b2 * GEOM_STRIDE;
          double Rn[9], Tn[3];
          load_node_rt(Rn, Tn, ga);
          mt_m(Ra, Rn, R); v_sub(Tt, T, Tn); mt_v(Ta, Rn, Tt);
          load_node_rt(Rn, Tn, gc);
          mt_m(Rc, Rn, R); v_sub(Tt, T, Tn); mt_v(Tc, Rn, Tt);
        }
        else
        {
          a1 = b1; a2 = mb.first_child; c1 = b1; c2 = a2 + 1;
          ga = gc = A.geom + (size_t)b1 * GEOM_STRIDE; gb_a = B.geom + (size_t)a2 * GEOM_STRIDE; gb_c = gb_a + GEOM_STRIDE;
          double Rn[9], Tn[3];
          load_node_rt(Rn, Tn, gb_a);
          m_m(Ra, R, Rn); m_v_p(Ta, R, Tn, T);
          load_node_rt(Rn, Tn, gb_c);
          m_m(Rc, R, Rn); m_v_p(Tc, R, Tn, T);
        }
        double d1 = rss_rect_dist(Ra, Ta, __ldg(ga + 12), __ldg(ga + 13), __ldg(gb_a + 12), __ldg(gb_a + 13), S);
        d1 -= (__ldg(ga + 14) + __ldg(gb_a + 14));
        d1 = (d1 < 0.0) ? 0.0 : d1;
        double d2 = rss_rect_dist(Rc, Tc, __ldg(gc + 12), __ldg(gc + 13), __ldg(gb_c + 12), __ldg(gb_c + 13), S);
        d2 -= (__ldg(gc + 14) + __ldg(gb_c + 14));
        d2 = (d2 < 0.0) ? 0.0 : d2;
        // res->distance = threshold, abs_err = rel_err = 0 (:1755-1760)
        const bool va = (d1 < (thr - 0)) || (d1 * (1.0 + 0) < thr), vc = (d2 < (thr - 0)) || (d2 * (1.0 + 0) < thr);
        const bool c_first = d2 < d1;
        // push the later one first
        for (int k = 0; k < 2; k++)
        {
          const bool push_c = (k == 0) ? !c_first : c_first;  // k = 0: the one visited second
          if (push_c ? vc : va)
          {
            const double *Rs = push_c ? Rc : Ra, *Ts = push_c ? Tc : Ta;
#pragma unroll
            for (int i = 0; i < 9; i++) STK(sp, i) = Rs[i];
            STK(sp, 9) = Ts[0]; STK(sp, 10) = Ts[1]; STK(sp, 11) = Ts[2];
            STK(sp, 12) = push_c ? __hiloint2double(c1, c2) : __hiloint2double(a1, a2);
            sp++;
          }
        }
      }
    }
    if (args.num_contact) args.num_contact[q] = count;
  }
#undef STK
}

}  // namespace c2a
