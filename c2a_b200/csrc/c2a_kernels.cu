// Batched controlled-conservative-advancement CCD on the GPU (sm_100a, FP64) + the C ABI of
// include/c2a_b200.h.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo.
//
// What runs on the device, per query (reference paths relative to /root/reference):
//   C2A_Solve's motion set-up                    C2A/src/C2A.cpp:2342-2395
//   C2A_QueryTimeOfContact  (the CA loop)        C2A/src/C2A.cpp:1987-2146
//   C2A_TimeOfContactStep   (per-step set-up)    C2A/src/C2A.cpp:1778-1931
//   TOCStepRecurse_Dis      (BVTT traversal)     C2A/src/C2A.cpp:1114-1354
//   pose outputs of C2A_Solve                    C2A/src/C2A.cpp:2411-2429
// with no host round trip between CA iterations.
//
// Execution model.  The traversal result is order dependent (res->distance shrinks as leaves are
// visited and gates pruning, C2A.cpp:1281-1351), so parallelism is taken ACROSS queries: one lane
// owns one query and commits node pairs in the reference's depth-first order.  The kernel is
// persistent: lanes claim queries from a global atomic counter until the batch is drained, so a
// lane whose query ended early (far-apart pair, one CA step) immediately starts another one.
// Every lane is a small state machine (ADVANCE / TRAVERSE / LEAF); the warp executes one phase at
// a time and votes (ballot) on which phase to run, so that lanes run the long FP64 routines
// (rectangle distance, triangle distance) together instead of serialising them against each other.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/c2a_b200.h"
#include "c2a_geom.cuh"
#include "c2a_motion.cuh"

namespace c2a {

// ---- device-resident model ---------------------------------------------------------------
// geom  [n][16]  R(9) Tr(3) l(2) r ang_radius      one 128-byte line per node; children adjacent
// rloc  [n][9]   R_loc (only read when a BV distance is non-zero)
// meta  [n]      {GetSize() = sqrt(l0^2+l1^2)+2r precomputed (PQP BV::GetSize), first_child}
// tris  [n][9]
constexpr int GEOM_STRIDE = 16;
struct NodeMeta { double size; int first_child; int pad; };
struct DevModel
{
  const double *geom;
  const double *rloc;
  const NodeMeta *meta;
  const double *tris;
  int n_nodes, n_tris;
};

constexpr int MAX_STACK = 64;     // >= depth(A)+depth(B)+2, validated on the host
constexpr int ENTRY_DOUBLES = 16; // R(9) T(3) d mint {b1,b2} pad  -> 128 B
constexpr int BLOCK_THREADS = 128;

struct BatchArgs
{
  DevModel A, B;
  const double *motions;  // [n][2][MOTION_DOUBLES], see c2a_motion.cuh
  const int *seedA, *seedB;
  long long n;
  double tol_d, tol_t;
  c2a_b200_results out;
  unsigned long long *counter;
};

enum LaneState { ST_ADVANCE = 0, ST_TRAVERSE = 1, ST_LEAF = 2, ST_EXIT = 3 };

C2A_DEV void load9(double d[9], const double *s)
{
#pragma unroll
  for (int i = 0; i < 9; i++) d[i] = __ldg(s + i);
}
C2A_DEV void load3(double d[3], const double *s) { d[0] = __ldg(s); d[1] = __ldg(s + 1); d[2] = __ldg(s + 2); }

__device__ __noinline__ double tri_distance_nl(const double R[9], const double T[3], const double *t1,
                                               const double *t2, double p[3], double q[3])
{
  double a[9], b[9];
  load9(a, t1);
  load9(b, t2);
  return tri_distance(R, T, a, b, p, q);
}

// One child BV test of an expansion (C2A.cpp:1237-1276): RSS distance, direction to world frame,
// the two directional motion bounds, and the child's conservative step bound.
//   gs: geom record of the side-1 node of the test, gt: of the side-2 node, rl: R_loc of the side-1 node.
C2A_DEV void child_test(const double Rc[9], const double Tc[3], const double *gs, const double *gt,
                        const double *rl, const double r1[9], const Motion &m1, const Motion &m2, double &d_out,
                        double &mint_out)
{
  double S[3];
  const double a0 = __ldg(gs + 12), a1 = __ldg(gs + 13), ra = __ldg(gs + 14);
  const double b0 = __ldg(gt + 12), b1 = __ldg(gt + 13), rb = __ldg(gt + 14);
  double d = rss_rect_dist(Rc, Tc, a0, a1, b0, b1, S);
  d -= (ra + rb);
  d = (d < 0.0) ? 0.0 : d;
  double mint = 0.0;
  if (d != 0.0)
  {
    double Rl[9], tmp[3], S1[3], S2[3];
    load9(Rl, rl);
    m_v(tmp, Rl, S);
    m_v(S1, r1, tmp);
    S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
    const double mb1 = motion_bound_bv(m1, __ldg(gs + 15), S1);
    const double mb2 = motion_bound_bv(m2, __ldg(gt + 15), S2);
    mint = (d) / (mb1 + mb2);
    if (mint <= 0) mint = 0.0;
  }
  d_out = d;
  mint_out = mint;
}

__global__ void __launch_bounds__(BLOCK_THREADS) c2a_solve_kernel(const BatchArgs args)
{
  const unsigned FULL = 0xffffffffu;
  const DevModel &A = args.A, &B = args.B;

  // per-lane traversal stack (local memory: interleaved across the warp by the hardware)
  double stk[MAX_STACK * ENTRY_DOUBLES];
  int sp = 0;

  // query state
  long long q = -1;
  Motion m1, m2;
  double r1[9], tt1[3];      // current pose of object 1 (objmotion1->transform)
  double Rrel[9], Trel[3];   // res->R, res->T
  double dist = 0, mint = 1, abs_err = 0, rel_err = 0, upbound = 1;
  double p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  double lamda = 0, lastLamda = 0;
  int numCA = 0, nItrs = 0, nbv = 0, ntri = 0;
  int seedA = 0, seedB = 0;
  int leaf_b1 = 0, leaf_b2 = 0;
  bool step_pending = false;  // a step set-up is due (first step or after advancing lamda)

  int state = ST_ADVANCE;

  while (true)
  {
    const unsigned mT = __ballot_sync(FULL, state == ST_TRAVERSE);
    const unsigned mL = __ballot_sync(FULL, state == ST_LEAF);
    const unsigned mA = __ballot_sync(FULL, state == ST_ADVANCE);
    if ((mT | mL | mA) == 0) break;
    const bool runL = (__popc(mL) >= 12) || (mT == 0 && mL != 0);
    const bool runA = (__popc(mA) >= 8) || (mT == 0 && !runL && mA != 0);

    // ------------------------------------------------------------------ ADVANCE ----------
    // claim a query / CA-loop bookkeeping after a finished step / next step's set-up
    if (runA && state == ST_ADVANCE)
    {
      bool finished = false, hit = false;
      if (q >= 0 && !step_pending)
      {
        // a step just ended: C2A_QueryTimeOfContact's loop, C2A.cpp:2053-2123
        if (numCA == 0) { numCA = 1; lastLamda = mint; }
        if (!(dist > args.tol_d)) { finished = true; hit = true; }
        else
        {
          nItrs++;
          if (nItrs > 150) { finished = true; hit = true; }
          else if (mint >= 1.0) { finished = true; hit = false; }
          else
          {
            const double dlamda = mint;
            if (dlamda < args.tol_t) { finished = true; hit = true; }
            else
            {
              lamda += dlamda;
              if (lamda >= 1.0) { finished = true; hit = false; }
              else
              {
                lastLamda = lamda;
                numCA++;
                motion_pose(m1, lamda, r1, tt1);
                upbound = 1.0 - lamda;
                step_pending = true;
              }
            }
          }
        }
        if (finished)
        {
          // C2A.cpp:2125-2143 and the pose outputs of C2A_Solve :2411-2429
          double toc = 0.0;
          const c2a_b200_results &o = args.out;
          if (hit)
          {
            toc = lastLamda;
            if (toc >= 1 - args.tol_t) toc = 0;
            if (o.pose_toc)
            {
              double R[9], T[3];
              motion_pose(m1, toc, R, T);
#pragma unroll
              for (int i = 0; i < 9; i++) o.pose_toc[24 * q + i] = R[i];
#pragma unroll
              for (int i = 0; i < 3; i++) o.pose_toc[24 * q + 9 + i] = T[i];
              motion_pose(m2, toc, R, T);
#pragma unroll
              for (int i = 0; i < 9; i++) o.pose_toc[24 * q + 12 + i] = R[i];
#pragma unroll
              for (int i = 0; i < 3; i++) o.pose_toc[24 * q + 21 + i] = T[i];
            }
          }
          if (o.status) o.status[q] = C2A_B200_QUERY_OK;
          if (o.collisionfree) o.collisionfree[q] = hit ? 0 : 1;
          if (o.num_ca) o.num_ca[q] = numCA;
          if (o.num_bv_tests) o.num_bv_tests[q] = nbv;
          if (o.num_tri_tests) o.num_tri_tests[q] = ntri;
          if (o.toc) o.toc[q] = toc;
          if (o.distance) o.distance[q] = dist;
          if (o.mint) o.mint[q] = mint;
          if (o.p1p2)
          {
#pragma unroll
            for (int i = 0; i < 3; i++) { o.p1p2[6 * q + i] = p1[i]; o.p1p2[6 * q + 3 + i] = p2[i]; }
          }
          q = -1;
        }
      }

      if (q < 0)
      {
        // claim the next query
        const long long nq = (long long)atomicAdd(args.counter, 1ull);
        if (nq >= args.n) state = ST_EXIT;
        else
        {
          q = nq;
          const double *pose = args.motions + (size_t)(2 * MOTION_DOUBLES) * q;
          motion_load(m1, pose);
          motion_load(m2, pose + MOTION_DOUBLES);
          seedA = args.seedA ? args.seedA[q] : 0;
          seedB = args.seedB ? args.seedB[q] : 0;
          if (m1.w < 1e-8 && m2.w < 1e-8)
          {
            // translation-only branch of the reference (C2A.cpp:2391-2395): not implemented
            if (args.out.status) args.out.status[q] = C2A_B200_QUERY_TRANSLATION_ONLY;
            q = -1;  // stay in ADVANCE: claim another one next round
          }
          else
          {
            load9(r1, pose);
            load3(tt1, pose + 9);
            numCA = 0; nItrs = 0; nbv = 0; ntri = 0;
            lamda = 0; lastLamda = 0; upbound = 1; mint = 1; dist = 0;
            p1[0] = p1[1] = p1[2] = p2[0] = p2[1] = p2[2] = 0;
            step_pending = true;
          }
        }
      }

      if (q >= 0 && step_pending)
      {
        // C2A_TimeOfContactStep, C2A.cpp:1791-1894
        double R2[9], T2[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3];
        if (numCA == 0)
        {
          const double *rec2 = args.motions + (size_t)(2 * MOTION_DOUBLES) * q + MOTION_DOUBLES;
          load9(R2, rec2); load3(T2, rec2 + 9);
        }
        else motion_pose(m2, lamda, R2, T2);
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

        double p[3], qq[3];
        dist = tri_distance_nl(Rrel, Trel, A.tris + 9 * seedA, B.tris + 9 * seedB, p, qq);
        if (numCA == 0) mint = 1;
        if (mint <= 0.005 || dist <= 0.5 || numCA > 5) { abs_err = 0; rel_err = 0; }
        else { abs_err = 1e+30; rel_err = (numCA <= 2) ? 3 : 0.5; }
        mint = 1;

        // root pair: always descended (d = -huge, mint = -1 pass every test)
        double *e = stk;
#pragma unroll
        for (int i = 0; i < 9; i++) e[i] = R[i];
        e[9] = T[0]; e[10] = T[1]; e[11] = T[2];
        e[12] = -1e300; e[13] = -1.0;
        e[14] = __hiloint2double(0, 0);
        sp = 1;
        step_pending = false;
        state = ST_TRAVERSE;
      }
    }

    // ------------------------------------------------------------------ TRAVERSE ---------
    if (state == ST_TRAVERSE)
    {
      // pop until an entry passes the descend test with the CURRENT distance (C2A.cpp:1281-1351);
      // entries that fail contribute their BV-level step bound
      int b1 = -1, b2 = -1;
      double R[9], T[3];
      while (sp > 0)
      {
        const double *e = stk + (sp - 1) * ENTRY_DOUBLES;
        sp--;
        const double d = e[12], mt = e[13];
        if (mt < upbound && ((d < (dist - abs_err)) || (d * (1 + rel_err) < dist)))
        {
          b1 = __double2hiint(e[14]); b2 = __double2loint(e[14]);
#pragma unroll
          for (int i = 0; i < 9; i++) R[i] = e[i];
          T[0] = e[9]; T[1] = e[10]; T[2] = e[11];
          break;
        }
        if (mt < mint) mint = mt;
      }
      if (b1 < 0) state = ST_ADVANCE;  // step finished
      else
      {
        const NodeMeta ma = A.meta[b1], mb = B.meta[b2];
        const bool l1 = ma.first_child < 0, l2 = mb.first_child < 0;
        if (l1 && l2) { leaf_b1 = b1; leaf_b2 = b2; state = ST_LEAF; }
        else
        {
          // expansion, C2A.cpp:1192-1279: two child pairs 'a' and 'c'
          int a1, a2, c1, c2;
          double Ra[9], Ta[3], Rc[9], Tc[3], d1, d2, mintb, minta;
          if (l2 || (!l1 && (ma.size > mb.size)))
          {
            a1 = ma.first_child; a2 = b2; c1 = a1 + 1; c2 = b2;
            const double *ga = A.geom + (size_t)a1 * GEOM_STRIDE, *gc = ga + GEOM_STRIDE;
            const double *gb = B.geom + (size_t)b2 * GEOM_STRIDE;
            double Rn[9], Tn[3], Tt[3];
            load9(Rn, ga); load3(Tn, ga + 9);
            mt_m(Ra, Rn, R); v_sub(Tt, T, Tn); mt_v(Ta, Rn, Tt);
            load9(Rn, gc); load3(Tn, gc + 9);
            mt_m(Rc, Rn, R); v_sub(Tt, T, Tn); mt_v(Tc, Rn, Tt);
            child_test(Ra, Ta, ga, gb, A.rloc + (size_t)a1 * 9, r1, m1, m2, d1, mintb);
            child_test(Rc, Tc, gc, gb, A.rloc + (size_t)c1 * 9, r1, m1, m2, d2, minta);
          }
          else
          {
            a1 = b1; a2 = mb.first_child; c1 = b1; c2 = a2 + 1;
            const double *ga = B.geom + (size_t)a2 * GEOM_STRIDE, *gc = ga + GEOM_STRIDE;
            const double *gb = A.geom + (size_t)b1 * GEOM_STRIDE;
            double Rn[9], Tn[3];
            load9(Rn, ga); load3(Tn, ga + 9);
            m_m(Ra, R, Rn); m_v_p(Ta, R, Tn, T);
            load9(Rn, gc); load3(Tn, gc + 9);
            m_m(Rc, R, Rn); m_v_p(Tc, R, Tn, T);
            child_test(Ra, Ta, gb, ga, A.rloc + (size_t)b1 * 9, r1, m1, m2, d1, mintb);
            child_test(Rc, Tc, gb, gc, A.rloc + (size_t)b1 * 9, r1, m1, m2, d2, minta);
          }
          nbv += 2;
          // push far child first, near child on top (visited first); ties visit 'a' first (d2 < d1 test)
          const bool c_first = d2 < d1;
          double *e0 = stk + sp * ENTRY_DOUBLES, *e1 = e0 + ENTRY_DOUBLES;
          double *ea = c_first ? e0 : e1, *ec = c_first ? e1 : e0;
#pragma unroll
          for (int i = 0; i < 9; i++) { ea[i] = Ra[i]; ec[i] = Rc[i]; }
#pragma unroll
          for (int i = 0; i < 3; i++) { ea[9 + i] = Ta[i]; ec[9 + i] = Tc[i]; }
          ea[12] = d1; ea[13] = mintb; ea[14] = __hiloint2double(a1, a2);
          ec[12] = d2; ec[13] = minta; ec[14] = __hiloint2double(c1, c2);
          sp += 2;
        }
      }
    }

    // ------------------------------------------------------------------ LEAF -------------
    if (runL && state == ST_LEAF)
    {
      // C2A.cpp:1141-1183
      double p[3], qq[3];
      const int ta = -A.meta[leaf_b1].first_child - 1, tb = -B.meta[leaf_b2].first_child - 1;
      const double dTri = tri_distance_nl(Rrel, Trel, A.tris + (size_t)9 * ta, B.tris + (size_t)9 * tb, p, qq);
      if (dTri <= dist)
      {
        dist = dTri;
        double w1[3], w2[3], S1[3], S2[3], tmp[3];
        m_v(tmp, r1, p); v_add(w1, tmp, tt1);
        m_v(tmp, r1, qq); v_add(w2, tmp, tt1);
        v_sub(S1, w2, w1);
        S2[0] = S1[0] * -1; S2[1] = S1[1] * -1; S2[2] = S1[2] * -1;
        v_cpy(p1, p); v_cpy(p2, qq);
        const double mb1 = motion_bound_leaf(m1, __ldg(A.geom + (size_t)leaf_b1 * GEOM_STRIDE + 15), S1);
        const double mb2 = motion_bound_leaf(m2, __ldg(B.geom + (size_t)leaf_b2 * GEOM_STRIDE + 15), S2);
        double mt = (dTri) / (mb1 + mb2);
        if (mt < 0.0) mt = 0.0;
        if (mt <= mint) mint = mt;
      }
      ntri++;
      state = ST_TRAVERSE;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const std::string &msg)
{
  g_err = msg;
  return code;
}
#define CUDA_TRY(x)                                                                                      \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) return fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

}  // namespace c2a

struct c2a_b200_model
{
  int device;
  int n_nodes, n_tris, depth;
  double *geom, *rloc, *tris;
  c2a::NodeMeta *meta;
};

using namespace c2a;

// ---- host side of the motion model ----------------------------------------------------------
// CInterpMotion ctor + CInterpMotion_Linear::velocity + LinearAngularVelocity
// (/root/reference/C2A/src/InterpMotion.cpp:148-168, 486-491, 228-270) for one object:
// pose = R0(9) T0(3) R1(9) T1(3)  ->  rec = R0(9) T0(3) cv(3) axis(3) w qs(4) pad.
// Runs on the host so that acos()/sqrt() are the host libm's, like the reference's.
static void motion_record_from_pose(const double *pose, double *rec)
{
  double qs[4], qt[4], q0[4], qd[4];
  for (int i = 0; i < 12; i++) rec[i] = pose[i];
  for (int i = 0; i < 3; i++) rec[12 + i] = pose[21 + i] - pose[9 + i];
  quat_from_matrix(qs, pose);
  quat_from_matrix(qt, pose + 12);
  q0[0] = -qs[0]; q0[1] = -qs[1]; q0[2] = -qs[2]; q0[3] = qs[3];
  quat_mul(qd, q0, qt);
  const double s = 1 < qd[3] ? 1 : qd[3];
  const double sign = s < 0 ? -1 : 1;
  const double a = (fabs(s - 1) <= 1e-40 || fabs(s + 1) <= 1e-40) ? (2 * sign)
                                                                  : (sign * acos(2 * s * s - 1) / sqrt(1 - s * s));
  const double tangent[3] = {a * qd[0], a * qd[1], a * qd[2]};
  rec[18] = sqrt(tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2]);
  const double len = tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2];
  if (len < (double)0.00000001f) { rec[15] = 1.0; rec[16] = 0.0; rec[17] = 0.0; }
  else
  {
    const double inv = 1.0 / sqrt(len);
    rec[15] = tangent[0] * inv; rec[16] = tangent[1] * inv; rec[17] = tangent[2] * inv;
  }
  rec[19] = qs[0]; rec[20] = qs[1]; rec[21] = qs[2]; rec[22] = qs[3];
  rec[23] = 0.0;
}

static void motions_from_poses_mt(const double *poses, int64_t n, double *motions, int n_threads)
{
  auto work = [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++)
    {
      motion_record_from_pose(poses + 48 * i, motions + 48 * i);
      motion_record_from_pose(poses + 48 * i + 24, motions + 48 * i + 24);
    }
  };
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  if (n_threads < 1) n_threads = 1;
  if ((int64_t)n_threads > n / 4096 + 1) n_threads = (int)(n / 4096 + 1);
  if (n_threads == 1) { work(0, n); return; }
  std::vector<std::thread> th;
  const int64_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; t++)
  {
    const int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto &t : th) t.join();
}

extern "C" {

const char *c2a_b200_last_error(void) { return g_err.c_str(); }
int64_t c2a_b200_launch_count(void) { return g_launches.load(); }

int c2a_b200_device_count(int32_t *count)
{
  if (!count) return fail(C2A_B200_ERR_ARG, "count is NULL");
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  *count = n;
  return C2A_B200_OK;
}

int c2a_b200_model_upload(const c2a_b200_bvh *bvh, int32_t device, c2a_b200_model **out)
{
  if (!bvh || !out) return fail(C2A_B200_ERR_ARG, "bvh/out is NULL");
  const int n = bvh->n_nodes, nt = bvh->n_tris;
  if (n <= 0 || nt <= 0 || !bvh->R || !bvh->Tr || !bvh->l || !bvh->r || !bvh->R_loc || !bvh->ang_radius ||
      !bvh->first_child || !bvh->tris)
    return fail(C2A_B200_ERR_ARG, "empty model or NULL array");

  // validate the topology and measure the depth (bounds the traversal stack)
  int depth = 0;
  {
    std::vector<std::pair<int, int>> todo;
    todo.push_back({0, 0});
    long long visited = 0;
    while (!todo.empty())
    {
      auto [i, d] = todo.back();
      todo.pop_back();
      if (++visited > n) return fail(C2A_B200_ERR_ARG, "BVH is not a tree");
      if (d > depth) depth = d;
      const int fc = bvh->first_child[i];
      if (fc < 0)
      {
        if (-fc - 1 >= nt) return fail(C2A_B200_ERR_ARG, "leaf triangle index out of range");
      }
      else
      {
        if (fc + 1 >= n || fc <= i) return fail(C2A_B200_ERR_ARG, "child index out of range");
        todo.push_back({fc, d + 1});
        todo.push_back({fc + 1, d + 1});
      }
    }
  }

  std::vector<double> geom((size_t)n * GEOM_STRIDE);
  std::vector<NodeMeta> meta(n);
  for (int i = 0; i < n; i++)
  {
    double *g = &geom[(size_t)i * GEOM_STRIDE];
    memcpy(g, bvh->R + 9 * (size_t)i, 9 * sizeof(double));
    memcpy(g + 9, bvh->Tr + 3 * (size_t)i, 3 * sizeof(double));
    g[12] = bvh->l[2 * (size_t)i]; g[13] = bvh->l[2 * (size_t)i + 1];
    g[14] = bvh->r[i];
    g[15] = bvh->ang_radius[i];
    // PQP BV::GetSize(), RSS form; sqrt/mul/add are correctly rounded on both sides, so precomputing is exact
    volatile double l0 = g[12] * g[12], l1 = g[13] * g[13];
    volatile double s = l0 + l1;
    meta[i].size = sqrt(s) + 2 * g[14];
    meta[i].first_child = bvh->first_child[i];
    meta[i].pad = 0;
  }

  CUDA_TRY(cudaSetDevice(device));
  c2a_b200_model *m = new c2a_b200_model();
  m->device = device; m->n_nodes = n; m->n_tris = nt; m->depth = depth;
  m->geom = m->rloc = m->tris = nullptr; m->meta = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&m->geom, geom.size() * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&m->rloc, (size_t)n * 9 * sizeof(double))) != cudaSuccess ||
      (e = cudaMalloc(&m->meta, (size_t)n * sizeof(NodeMeta))) != cudaSuccess ||
      (e = cudaMalloc(&m->tris, (size_t)nt * 9 * sizeof(double))) != cudaSuccess ||
      (e = cudaMemcpy(m->geom, geom.data(), geom.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(m->rloc, bvh->R_loc, (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(m->meta, meta.data(), (size_t)n * sizeof(NodeMeta), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(m->tris, bvh->tris, (size_t)nt * 9 * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
  {
    c2a_b200_model_free(m);
    return fail(C2A_B200_ERR_CUDA, std::string("model upload: ") + cudaGetErrorString(e));
  }
  *out = m;
  return C2A_B200_OK;
}

int c2a_b200_model_free(c2a_b200_model *m)
{
  if (!m) return C2A_B200_OK;
  cudaSetDevice(m->device);
  cudaFree(m->geom); cudaFree(m->rloc); cudaFree(m->meta); cudaFree(m->tris);
  delete m;
  return C2A_B200_OK;
}

int c2a_b200_model_info(const c2a_b200_model *m, int32_t *device, int32_t *n_nodes, int32_t *n_tris, int32_t *depth)
{
  if (!m) return fail(C2A_B200_ERR_ARG, "model is NULL");
  if (device) *device = m->device;
  if (n_nodes) *n_nodes = m->n_nodes;
  if (n_tris) *n_tris = m->n_tris;
  if (depth) *depth = m->depth;
  return C2A_B200_OK;
}

// per-device scratch: the claim counter
static int launch_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses, const int32_t *sa,
                        const int32_t *sb, int64_t n, double tol_d, double tol_t, const c2a_b200_results *out,
                        unsigned long long *counter, cudaStream_t stream)
{
  BatchArgs args;
  args.A = DevModel{a->geom, a->rloc, a->meta, a->tris, a->n_nodes, a->n_tris};
  args.B = DevModel{b->geom, b->rloc, b->meta, b->tris, b->n_nodes, b->n_tris};
  args.motions = poses; args.seedA = sa; args.seedB = sb; args.n = n;
  args.tol_d = tol_d; args.tol_t = tol_t; args.out = *out; args.counter = counter;

  int sms = 0, per_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, a->device));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c2a_solve_kernel, BLOCK_THREADS, 0));
  if (per_sm < 1) per_sm = 1;
  long long blocks = (long long)sms * per_sm;  // persistent: one resident wave
  const long long need = (n + BLOCK_THREADS - 1) / BLOCK_THREADS;
  if (blocks > need) blocks = need;
  if (blocks < 1) blocks = 1;
  CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream));
  c2a_solve_kernel<<<(unsigned)blocks, BLOCK_THREADS, 0, stream>>>(args);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaGetLastError());
  return C2A_B200_OK;
}

static int check_pair(const c2a_b200_model *a, const c2a_b200_model *b, int64_t n, const void *poses,
                      const c2a_b200_results *out)
{
  if (!a || !b || !out || (n > 0 && !poses)) return fail(C2A_B200_ERR_ARG, "NULL argument");
  if (n < 0) return fail(C2A_B200_ERR_ARG, "negative batch size");
  if (a->device != b->device) return fail(C2A_B200_ERR_DEVICE, "models live on different devices");
  if (a->depth + b->depth + 2 > MAX_STACK)
    return fail(C2A_B200_ERR_DEPTH, "BVH depths exceed the traversal stack (" + std::to_string(a->depth) + "+" +
                                        std::to_string(b->depth) + ")");
  return C2A_B200_OK;
}

int c2a_b200_motions_from_poses(const double *poses, int64_t n, double *motions, int32_t n_threads)
{
  if (n < 0 || (n > 0 && (!poses || !motions))) return fail(C2A_B200_ERR_ARG, "NULL argument");
  motions_from_poses_mt(poses, n, motions, n_threads);
  return C2A_B200_OK;
}

int c2a_b200_solve_batch_device(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses_dev,
                                const int32_t *seed_a_dev, const int32_t *seed_b_dev, int64_t n, double tol_d,
                                double tol_t, const c2a_b200_results *out_dev, void *cuda_stream)
{
  int rc = check_pair(a, b, n, poses_dev, out_dev);
  if (rc) return rc;
  if (n == 0) return C2A_B200_OK;
  CUDA_TRY(cudaSetDevice(a->device));
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  unsigned long long *counter = nullptr;
  CUDA_TRY(cudaMallocAsync(&counter, sizeof(unsigned long long), stream));
  rc = launch_batch(a, b, poses_dev, seed_a_dev, seed_b_dev, n, tol_d, tol_t, out_dev, counter, stream);
  cudaFreeAsync(counter, stream);
  return rc;
}

int c2a_b200_solve_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses,
                         const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d, double tol_t,
                         const c2a_b200_results *out)
{
  int rc = check_pair(a, b, n, poses, out);
  if (rc) return rc;
  if (n == 0) return C2A_B200_OK;
  CUDA_TRY(cudaSetDevice(a->device));
  cudaStream_t stream;
  CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));

  // one device arena: poses | seeds | outputs | counter
  const size_t N = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_pose = take(N * 48 * 8);
  const size_t o_sa = seed_a ? take(N * 4) : 0, o_sb = seed_b ? take(N * 4) : 0;
  const size_t o_status = out->status ? take(N * 4) : 0, o_cf = out->collisionfree ? take(N * 4) : 0;
  const size_t o_nca = out->num_ca ? take(N * 4) : 0, o_nbv = out->num_bv_tests ? take(N * 4) : 0;
  const size_t o_ntri = out->num_tri_tests ? take(N * 4) : 0;
  const size_t o_toc = out->toc ? take(N * 8) : 0, o_dist = out->distance ? take(N * 8) : 0;
  const size_t o_mint = out->mint ? take(N * 8) : 0, o_pp = out->p1p2 ? take(N * 48) : 0;
  const size_t o_pt = out->pose_toc ? take(N * 192) : 0;
  const size_t o_cnt = take(8);
  char *arena = nullptr;
  cudaError_t e = cudaMallocAsync(&arena, off, stream);
  if (e != cudaSuccess)
  {
    cudaStreamDestroy(stream);
    return fail(C2A_B200_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e));
  }
  c2a_b200_results d;
  memset(&d, 0, sizeof(d));
  if (out->status) d.status = (int32_t *)(arena + o_status);
  if (out->collisionfree) d.collisionfree = (int32_t *)(arena + o_cf);
  if (out->num_ca) d.num_ca = (int32_t *)(arena + o_nca);
  if (out->num_bv_tests) d.num_bv_tests = (int32_t *)(arena + o_nbv);
  if (out->num_tri_tests) d.num_tri_tests = (int32_t *)(arena + o_ntri);
  if (out->toc) d.toc = (double *)(arena + o_toc);
  if (out->distance) d.distance = (double *)(arena + o_dist);
  if (out->mint) d.mint = (double *)(arena + o_mint);
  if (out->p1p2) d.p1p2 = (double *)(arena + o_pp);
  if (out->pose_toc) d.pose_toc = (double *)(arena + o_pt);

  rc = C2A_B200_OK;
#define STEP(x)                                                                              \
  if (rc == C2A_B200_OK && (e = (x)) != cudaSuccess) rc = fail(C2A_B200_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e));
  STEP(cudaMemsetAsync(arena, 0, off, stream));
  // motion constants on the host (libm acos), straight into pinned staging memory
  double *staging = nullptr;
  STEP(cudaMallocHost(&staging, N * 48 * 8));
  if (rc == C2A_B200_OK) motions_from_poses_mt(poses, n, staging, 0);
  STEP(cudaMemcpyAsync(arena + o_pose, staging, N * 48 * 8, cudaMemcpyHostToDevice, stream));
  if (seed_a) STEP(cudaMemcpyAsync(arena + o_sa, seed_a, N * 4, cudaMemcpyHostToDevice, stream));
  if (seed_b) STEP(cudaMemcpyAsync(arena + o_sb, seed_b, N * 4, cudaMemcpyHostToDevice, stream));
  if (out->pose_toc) STEP(cudaMemsetAsync(arena + o_pt, 0, N * 192, stream));
  if (rc == C2A_B200_OK)
    rc = launch_batch(a, b, (const double *)(arena + o_pose), seed_a ? (const int32_t *)(arena + o_sa) : nullptr,
                      seed_b ? (const int32_t *)(arena + o_sb) : nullptr, n, tol_d, tol_t, &d,
                      (unsigned long long *)(arena + o_cnt), stream);
#define BACK(field, ofs, bytes) \
  if (out->field) STEP(cudaMemcpyAsync(out->field, arena + ofs, bytes, cudaMemcpyDeviceToHost, stream));
  BACK(status, o_status, N * 4) BACK(collisionfree, o_cf, N * 4) BACK(num_ca, o_nca, N * 4)
  BACK(num_bv_tests, o_nbv, N * 4) BACK(num_tri_tests, o_ntri, N * 4) BACK(toc, o_toc, N * 8)
  BACK(distance, o_dist, N * 8) BACK(mint, o_mint, N * 8) BACK(p1p2, o_pp, N * 48) BACK(pose_toc, o_pt, N * 192)
#undef BACK
  STEP(cudaStreamSynchronize(stream));
#undef STEP
  cudaFreeAsync(arena, stream);
  cudaStreamSynchronize(stream);
  if (staging) cudaFreeHost(staging);
  cudaStreamDestroy(stream);
  return rc;
}

}  // extern "C"

// ---- unit-test hooks (include/c2a_b200_testing.h) -------------------------------------------
#include "../../include/c2a_b200_testing.h"

namespace c2a {
__global__ void k_test_rect_dist(const double *R, const double *T, const double *ab, long long n, double *dist, double *S)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[9], t[3], s[3] = {S[3 * i], S[3 * i + 1], S[3 * i + 2]};
  for (int k = 0; k < 9; k++) r[k] = R[9 * i + k];
  for (int k = 0; k < 3; k++) t[k] = T[3 * i + k];
  dist[i] = rss_rect_dist(r, t, ab[4 * i], ab[4 * i + 1], ab[4 * i + 2], ab[4 * i + 3], s);
  S[3 * i] = s[0]; S[3 * i + 1] = s[1]; S[3 * i + 2] = s[2];
}
__global__ void k_test_tri_distance(const double *R, const double *T, const double *t1, const double *t2, long long n,
                                    double *dist, double *pq)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[9], t[3], p[3], q[3];
  for (int k = 0; k < 9; k++) r[k] = R[9 * i + k];
  for (int k = 0; k < 3; k++) t[k] = T[3 * i + k];
  dist[i] = tri_distance_nl(r, t, t1 + 9 * i, t2 + 9 * i, p, q);
  for (int k = 0; k < 3; k++) { pq[6 * i + k] = p[k]; pq[6 * i + 3 + k] = q[k]; }
}
__global__ void k_test_motion(const double *rec, const double *t, const double *ar, const double *dir, long long n, double *out)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Motion m;
  motion_load(m, rec + 24 * i);
  double R[9], T[3];
  motion_pose(m, t[i], R, T);
  for (int k = 0; k < 9; k++) out[14 * i + k] = R[k];
  for (int k = 0; k < 3; k++) out[14 * i + 9 + k] = T[k];
  double n1[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]}, n2[3] = {dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
  out[14 * i + 12] = motion_bound_bv(m, ar[i], n1);
  out[14 * i + 13] = motion_bound_leaf(m, ar[i], n2);
}
__global__ void k_test_sincos(const double *x, long long n, double *s, double *c)
{
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  s[i] = libm_sin(x[i]);
  c[i] = libm_cos(x[i]);
}

struct DevBuf
{
  std::vector<void *> ptrs;
  ~DevBuf() { for (void *p : ptrs) cudaFree(p); }
  template <class T> cudaError_t up(T **d, const T *h, size_t count)
  {
    cudaError_t e = cudaMalloc((void **)d, count * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(*d);
    return h ? cudaMemcpy(*d, h, count * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};
}  // namespace c2a

extern "C" {

int c2a_b200_test_rect_dist(const double *R, const double *T, const double *ab, int64_t n, double *dist, double *S)
{
  DevBuf b; double *dR, *dT, *dab, *dd, *dS;
  CUDA_TRY(b.up(&dR, R, 9 * n)); CUDA_TRY(b.up(&dT, T, 3 * n)); CUDA_TRY(b.up(&dab, ab, 4 * n));
  CUDA_TRY(b.up(&dd, (const double *)nullptr, n)); CUDA_TRY(b.up(&dS, S, 3 * n));
  k_test_rect_dist<<<(unsigned)((n + 127) / 128), 128>>>(dR, dT, dab, n, dd, dS);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(dist, dd, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(S, dS, 3 * n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_test_tri_distance(const double *R, const double *T, const double *t1, const double *t2, int64_t n,
                               double *dist, double *pq)
{
  DevBuf b; double *dR, *dT, *d1, *d2, *dd, *dpq;
  CUDA_TRY(b.up(&dR, R, 9 * n)); CUDA_TRY(b.up(&dT, T, 3 * n)); CUDA_TRY(b.up(&d1, t1, 9 * n)); CUDA_TRY(b.up(&d2, t2, 9 * n));
  CUDA_TRY(b.up(&dd, (const double *)nullptr, n)); CUDA_TRY(b.up(&dpq, (const double *)nullptr, 6 * n));
  k_test_tri_distance<<<(unsigned)((n + 127) / 128), 128>>>(dR, dT, d1, d2, n, dd, dpq);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(dist, dd, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(pq, dpq, 6 * n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_test_motion(const double *rec, const double *t, const double *ang_radius, const double *dir, int64_t n,
                         double *out)
{
  DevBuf b; double *dr, *dt, *da, *dd, *dout;
  CUDA_TRY(b.up(&dr, rec, 24 * n)); CUDA_TRY(b.up(&dt, t, n)); CUDA_TRY(b.up(&da, ang_radius, n)); CUDA_TRY(b.up(&dd, dir, 3 * n));
  CUDA_TRY(b.up(&dout, (const double *)nullptr, 14 * n));
  k_test_motion<<<(unsigned)((n + 127) / 128), 128>>>(dr, dt, da, dd, n, dout);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, dout, 14 * n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_test_sincos(const double *x, int64_t n, double *s, double *c)
{
  DevBuf b; double *dx, *ds, *dc;
  CUDA_TRY(b.up(&dx, x, n)); CUDA_TRY(b.up(&ds, (const double *)nullptr, n)); CUDA_TRY(b.up(&dc, (const double *)nullptr, n));
  k_test_sincos<<<(unsigned)((n + 127) / 128), 128>>>(dx, n, ds, dc);
  g_launches.fetch_add(1);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(s, ds, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(c, dc, n * 8, cudaMemcpyDeviceToHost));
  return C2A_B200_OK;
}

int c2a_b200_host_sincos(const double *x, int64_t n, double *s, double *c)
{
  for (int64_t i = 0; i < n; i++) { s[i] = libm_sin(x[i]); c[i] = libm_cos(x[i]); }
  return C2A_B200_OK;
}

}  // extern "C"
