// Wide traversal kernel: one warp advances ONE query, popping up to 16 node pairs per round (sm_100a, FP64).
//
// c2a_solve_kernel (c2a_solve.cuh) takes its parallelism across queries: a query alone on its warp -- the tail of a
// launch, a small batch, a single C2A_Solve call -- commits 2.85 expansions per 13.7 k-cycle pass.  This kernel takes
// over such queries at a CA-step boundary (a 64-byte record, written by c2a_solve_kernel once the claim queue is
// empty and the warp is running out of queries) and runs their remaining CA steps with parallelism INSIDE the
// traversal, bit-identical to the reference's depth-first order (C2A/src/C2A.cpp:1114-1354):
//
//   * the pending node pairs live on a per-warp stack in global memory, in the reference's visiting order (top =
//     visited next).  A round pops the top 16 pairs, runs their 32 child tests side by side (a lane per test) and
//     pushes the children back in visiting order -- of one pair: the closer child on top (C2A.cpp:1281: d2 < d1);
//     of the window: the first pair's children on top -- so the stack stays ordered without sorting.  Leaf pairs go
//     to a waiting list instead and are tested 32 at a time (LEAF pass);
//   * every node carries a pre-order key (one bit per level: 0 = the child visited first) and M = the maximum over
//     its path of the test values d (+inf where mint >= UpboundTOC);
//   * the running distance changes only at leaves with dTri <= dist (C2A.cpp:1150: "events").  Everything that
//     precedes the top of the stack and the waiting leaves has been evaluated, so the events up to there are
//     resolved exactly, in key order: Dw, the distance in force at the top of the stack.  A pair is expanded and its
//     children are pushed iff M < Dw: the distance only shrinks, so this is a superset of what the reference
//     visits; the pairs of one window are expanded without regard to the events inside the window;
//   * when nothing is pending all events are known.  FOLD over the child records {key, M of the parent, own value,
//     own step bound}: with D(key) the distance in force at a node's pre-order position, a node was visited iff
//     M < D(key); visited parents count their child tests (num_bv_tests), children that fail their own test at
//     their own position fold their step bound (res->mint, C2A.cpp:1292-1350), visited leaves count a triangle test;
//     event leaves fold theirs and the last one gives distance, p1/p2 and last_tri.
//
// "visited <=> M(n) < D(key(n))": => is immediate (D never grows).  <= could only fail if an ancestor's test value
// of an event leaf l is >= dTri(l) (that edge passed under an earlier, larger distance; descendants visited after l
// would be mis-classified).  A bounding-volume distance never exceeds the distance of the triangles inside up to
// rounding; the case is DETECTED (M of the leaf's parent >= dTri != 0) and the step is then redone by the same
// code popping ONE pair per round with direct bookkeeping -- the reference's walk itself.  The CPU statement of
// the scheme is oracle/c2a_oracle.cpp wide_step (bit-identical to the sequential port on every fixture; no such
// case in 21 k steps / 1 M events).  Only exact-mode steps (abs_err = rel_err = 0: every step past a query's
// fifth, C2A.cpp:1869) run wide; the others run one pair per round.
#pragma once
#include "c2a_solve.cuh"

// The counters of c2a_b200_wide_stats (cycles per phase, rounds, tests) cost ~15 % of the kernel's time and skew what
// they measure (clock reads and atomics between the passes), so they are compiled out of the product build;
// scripts/build_variant.py <name> -DC2A_WIDE_STATS=1 makes a library with them.
#ifndef C2A_WIDE_STATS
#define C2A_WIDE_STATS 0
#endif

namespace c2a {

constexpr int WIDE_WPB = 4;                 // warps per block
constexpr int WIDE_THREADS = 32 * WIDE_WPB;
constexpr int WIDE_PL = 96;                 // waiting leaf pairs (never more than 31 + 32)
constexpr int WIDE_UL = 128;                // evaluated leaves whose event status is open
constexpr int WIDE_EV = 768;                // events per step
constexpr int WIDE_CST = 40;                // step constants (doubles)
constexpr int WIDE_RND = 1024;              // rounds of a step whose {first record, events resolved so far} are kept for the fold
constexpr int WIDE_KEY_LEVELS = 56;         // path bits in a key (bits 63..8); bit 7 = leaf pair, bits 5..0 = depth
// shared memory per warp: step constants, waiting leaves, open leaves -- kept small (7.5 KB) on purpose: the node and
// stack fetches live in L1, and every KB of shared memory is a KB less of it (with the events and the round table in
// shared memory as well, 24 KB per warp, a round took 27 k cycles instead of 14.5 k)
constexpr size_t WIDE_WARP_SMEM = (size_t)WIDE_CST * 8 + (size_t)WIDE_PL * (8 + 8 + 8 + 4 + 4) + (size_t)WIDE_UL * 32;
constexpr size_t WIDE_AUX_DOUBLES = (size_t)WIDE_EV * 2 + (size_t)WIDE_RND;   // per warp, global: events {key, d}, round table {rec0, nev}
constexpr size_t WIDE_BLOCK_SMEM = WIDE_WARP_SMEM * WIDE_WPB;
constexpr int WIDE_LEAFOUT_DOUBLES = 8;     // p(3) q(3) leaf step bound, {ta, tb}

struct WideArgs
{
  DevModel A, B;
  const double *motions;
  const int *seedA, *seedB;
  double tol_d, tol_t;
  c2a_b200_results out;
  const double *items;                  // [SPILL_BUCKETS][items_cap][MB_DOUBLES]: CA-loop state of the queries handed over
  unsigned long long *ctl;              // the main kernel's control words (SPILL_CTL_*): list lengths, claim counters, progress
  long long items_cap;
  int early;                            // launched beside the main kernel (side stream): lists still growing, poll until it is done
  unsigned main_blocks;                 // grid of the main kernel
  double *stack;                        // [warps][stack_cap][ENTRY_DOUBLES]
  double *recs;                         // [warps][rec_cap][4]
  double *leafout;                      // [warps][WIDE_UL][WIDE_LEAFOUT_DOUBLES]
  double *aux;                          // [warps][WIDE_AUX_DOUBLES]
  int stack_cap, rec_cap;
  int window;                           // pairs per round (1..16)
  unsigned long long *stats;            // optional [WIDE_NSTATS]
  unsigned long long *trace;
};
enum { WS_STEPS = 0, WS_REDO, WS_ROUNDS, WS_LEAF_PASSES, WS_TESTS, WS_LEAVES, WS_EVENTS, WS_CYC_EXPAND, WS_CYC_LEAF,
       WS_CYC_RESOLVE, WS_CYC_FOLD, WS_CYC_SETUP, WS_QUERIES, WS_SEQ_STEPS, WS_T_FIRST, WS_T_LAST,
       WS_X_SELECT, WS_X_POP, WS_X_XFORM, WS_X_RECT, WS_X_BOUND, WS_X_SYNC, WS_X_RECORD, WS_X_PUSH, WIDE_NSTATS };  // WS_X_*: lane 0's cycles inside an EXPAND pass
#if C2A_WIDE_STATS
#define WIDE_TS(var, dep) { asm volatile("" :: "d"((double)(dep)) : "memory"); var = clock64(); }
#else
#define WIDE_TS(var, dep)
#endif

C2A_DEV unsigned long long shfl_u64(unsigned mask, unsigned long long v, int src) { return __shfl_sync(mask, v, src); }
C2A_DEV unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1)
  {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

__global__ void __launch_bounds__(WIDE_THREADS, 2) c2a_wide_kernel(const WideArgs args)
{
  extern __shared__ double smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * WIDE_WPB + warp;
  const DevModel &A = args.A, &B = args.B;
  const double INF = __longlong_as_double(0x7ff0000000000000ll);

  // ---- per-warp shared memory
  char *base = reinterpret_cast<char *>(smem) + (size_t)warp * WIDE_WARP_SMEM;
  double *cst = reinterpret_cast<double *>(base);
  unsigned long long *pl_key = reinterpret_cast<unsigned long long *>(cst + WIDE_CST);
  double *pl_mpar = reinterpret_cast<double *>(pl_key + WIDE_PL);
  double *pl_val = pl_mpar + WIDE_PL;
  int *pl_b1 = reinterpret_cast<int *>(pl_val + WIDE_PL);
  int *pl_b2 = pl_b1 + WIDE_PL;
  unsigned long long *ul_key = reinterpret_cast<unsigned long long *>(pl_b2 + WIDE_PL);
  double *ul_mpar = reinterpret_cast<double *>(ul_key + WIDE_UL);
  double *ul_val = ul_mpar + WIDE_UL;
  double *ul_dtri = ul_val + WIDE_UL;
  // events and round table: per-warp global memory (written and read by this warp only; L1/L2 resident)
  double *const aux = args.aux + (size_t)gw * WIDE_AUX_DOUBLES;
  unsigned long long *ev_key = reinterpret_cast<unsigned long long *>(aux);
  double *ev_d = aux + WIDE_EV;
  int2 *rnd_tab = reinterpret_cast<int2 *>(aux + 2 * WIDE_EV);   // {first record of round r, events resolved when it started}
  enum { C_R1 = 0, C_TT1 = 9, C_CV1 = 12, C_AX1 = 15, C_W1 = 18, C_CV2 = 19, C_AX2 = 22, C_W2 = 25, C_RREL = 26, C_TREL = 35, C_X = 38 };

  double *const stk = args.stack + (size_t)gw * args.stack_cap * ENTRY_DOUBLES;
  double *const recs = args.recs + (size_t)gw * args.rec_cap * 4;
  double *const leafout = args.leafout + (size_t)gw * WIDE_UL * WIDE_LEAFOUT_DOUBLES;

  // The early launch only keeps an SM slot while EVERY block of the main kernel is already on the machine (it waits for
  // that kernel to finish, so it must never be what stands between one of its blocks and an SM); the launch after the
  // main kernel picks up whatever this one left.
  if (args.early && *(volatile unsigned long long *)(args.ctl + SPILL_CTL_STARTED) < (unsigned long long)args.main_blocks) return;
  if ((C2A_WIDE_STATS && args.stats) && threadIdx.x == 0) atomicMin(args.stats + WS_T_FIRST, global_ns());
  while (true)
  {
    // ---- claim a handed-over query: the lists in order (most CA steps taken so far first).  In the early launch the
    // main kernel is still appending: wait for more until all its warps are done
    int bucket = -1;
    unsigned long long bi = 0;
    if (lane == 0)
    {
      volatile unsigned long long *ctl = args.ctl;
      bool final_scan = !args.early;
      while (true)
      {
        const bool main_done = final_scan || ctl[SPILL_CTL_DONE] >= (unsigned long long)args.main_blocks * WARPS_PER_BLOCK;
        for (int b = 0; b < SPILL_BUCKETS && bucket < 0; b++)
        {
          unsigned long long c = ctl[SPILL_CTL_CLAIMED + b];
          while (c < min(ctl[b], (unsigned long long)args.items_cap))
          {
            const unsigned long long old = atomicCAS(args.ctl + SPILL_CTL_CLAIMED + b, c, c + 1);
            if (old == c) { bucket = b; bi = c; break; }
            c = old;
          }
        }
        if (bucket >= 0 || final_scan) break;
        if (main_done) final_scan = true;   // the lengths read after this point are final: one more look
        else __nanosleep(2000);
      }
    }
    bucket = __shfl_sync(FULL, bucket, 0);
    bi = shfl_u64(FULL, bi, 0);
    if (bucket < 0) break;
    const double *r = args.items + ((size_t)bucket * args.items_cap + bi) * MB_DOUBLES;
    long long q;
    while ((q = *reinterpret_cast<const volatile long long *>(r)) < 0) __nanosleep(200);   // reserved, not yet written
    __threadfence();
    double lamda = __ldcg(r + 1), lastLamda = __ldcg(r + 2), mint = __ldcg(r + 3), upb = __ldcg(r + 4);
    const double c5 = __ldcg(r + 5), c6 = __ldcg(r + 6), c7 = __ldcg(r + 7);
    int numCA = __double2hiint(c5), nItrs = __double2loint(c5);
    int nbv = __double2hiint(c6), ntri = __double2loint(c6);
    int lastA = __double2hiint(c7), lastB = __double2loint(c7);
    const double *rec = args.motions + (size_t)(2 * MOTION_DOUBLES) * q;
    const int seedA = seed_or_zero(args.seedA, q, A.n_tris), seedB = seed_or_zero(args.seedB, q, B.n_tris);
    double dist = 0;
    if ((C2A_WIDE_STATS && args.stats) && lane == 0) atomicAdd(args.stats + WS_QUERIES, 1ull);

    while (true)
    {
      // ================================================================ one CA step =========
      // C2A_TimeOfContactStep, C2A.cpp:1791-1894 (lane 0; the constants go to shared memory)
      const long long t_setup = (C2A_WIDE_STATS && args.stats) ? clock64() : 0;
      int root_leaf = 0;
      if (lane == 0)
      {
        double r1[9], tt1[3], R2[9], T2[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3], Rrel[9], Trel[3];
        if (numCA == 0)
        {
          load9(r1, rec); load3(tt1, rec + 9);
          load9(R2, rec + MOTION_DOUBLES); load3(T2, rec + MOTION_DOUBLES + 9);
        }
        else
        {
          motion_pose_nl(rec, lamda, r1, tt1);
          motion_pose_nl(rec + MOTION_DOUBLES, lamda, R2, T2);
        }
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
        const double sd = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * seedA, B.tris + (size_t)TRI_STRIDE * seedB, p, qq);
#pragma unroll
        for (int i = 0; i < 9; i++) { cst[C_R1 + i] = r1[i]; cst[C_RREL + i] = Rrel[i]; }
#pragma unroll
        for (int i = 0; i < 3; i++)
        {
          cst[C_TT1 + i] = tt1[i]; cst[C_TREL + i] = Trel[i];
          cst[C_CV1 + i] = __ldg(rec + 12 + i); cst[C_AX1 + i] = __ldg(rec + 15 + i);
          cst[C_CV2 + i] = __ldg(rec + MOTION_DOUBLES + 12 + i); cst[C_AX2 + i] = __ldg(rec + MOTION_DOUBLES + 15 + i);
        }
        cst[C_W1] = __ldg(rec + 18); cst[C_W2] = __ldg(rec + MOTION_DOUBLES + 18);
        cst[C_X] = sd;
        // the root pair is descended unconditionally: stack entry 0
        double2 *e2 = reinterpret_cast<double2 *>(stk);
        e2[0] = make_double2(R[0], R[1]); e2[1] = make_double2(R[2], R[3]);
        e2[2] = make_double2(R[4], R[5]); e2[3] = make_double2(R[6], R[7]);
        e2[4] = make_double2(R[8], T[0]); e2[5] = make_double2(T[1], T[2]);
        e2[6] = make_double2(-INF, __hiloint2double(0, 0)); e2[7] = make_double2(__longlong_as_double(0ll), 0.0);
        root_leaf = (A.meta[0].first_child < 0 && B.meta[0].first_child < 0) ? 1 : 0;
      }
      __syncwarp();
      root_leaf = __shfl_sync(FULL, root_leaf, 0);
      const double seed_dist = cst[C_X];
      double mint_prev = mint;
      if (numCA == 0) mint_prev = 1;
      double abs_err, rel_err;
      if (mint_prev <= 0.005 || seed_dist <= 0.5 || numCA > 5) { abs_err = 0; rel_err = 0; }
      else { abs_err = 1e+30; rel_err = (numCA <= 2) ? 3 : 0.5; }
      const bool exact = abs_err == 0 && rel_err == 0;
      if ((C2A_WIDE_STATS && args.stats) && lane == 0) { atomicAdd(args.stats + WS_STEPS, 1ull); atomicAdd(args.stats + WS_CYC_SETUP, (unsigned long long)(clock64() - t_setup)); }

      // ---- traversal.  seq = false: wide rounds with records, events and fold; seq = true: one pair per round with
      // direct bookkeeping (non-exact steps, and the redo after an anomaly / a full arena)
      bool seq = !exact || args.window <= 1;
      double step_mint = 1, step_dist = seed_dist;
      int step_nbv = 0, step_ntri = 0, step_lastA = lastA, step_lastB = lastB;
      bool have_best = false;
      double best_p[3] = {0, 0, 0}, best_q[3] = {0, 0, 0};
      while (true)
      {
        // (re)start the step's traversal
        int sp = 1, nrec = 0, npl = 0, nev = 0, nrnd = 0;
        unsigned ulm0 = 0, ulm1 = 0, ulm2 = 0, ulm3 = 0;  // occupied slots of the open-leaf list
        double Dw = seed_dist;
        bool redo = false;
        step_mint = 1; step_nbv = 0; step_ntri = 0; step_lastA = lastA; step_lastB = lastB; have_best = false;
        if (root_leaf)
        {
          // degenerate: both models are single triangles -- the root pair is a leaf pair
          sp = 0;
          if (lane == 0) { pl_key[0] = 0x80ull; pl_mpar[0] = -INF; pl_val[0] = -INF; pl_b1[0] = 0; pl_b2[0] = 0; }
          npl = 1;
          __syncwarp();
        }
        const int Wn = seq ? 1 : args.window;

        while (sp > 0 || npl > 0)
        {
          const long long t_pass = (C2A_WIDE_STATS && args.stats) ? clock64() : 0;
          bool did_leaf = false;
          unsigned long long top_key = ~0ull;   // key of the new top of the stack if this round pushed
          bool top_known = false;
          if (npl >= 32 || sp == 0 || (seq && npl > 0))
          {
            // ------------------------------------------------------------ LEAF pass (C2A.cpp:1141-1183)
            did_leaf = true;
            const int n_take = npl < 32 ? npl : 32;
            unsigned long long key = 0; double mpar = 0, val = 0; int b1 = 0, b2 = 0;
            bool go = false;
            if (lane < n_take)
            {
              key = pl_key[lane]; mpar = pl_mpar[lane]; val = pl_val[lane]; b1 = pl_b1[lane]; b2 = pl_b2[lane];
              const double M = mpar > val ? mpar : val;
              go = seq ? true : (M < Dw);
            }
            double dTri = 0, leaf_mt = 0, p[3] = {0, 0, 0}, qq[3] = {0, 0, 0};
            int ta = 0, tb = 0;
            if (go)
            {
              double Rrel[9], Trel[3];
#pragma unroll
              for (int i = 0; i < 9; i++) Rrel[i] = cst[C_RREL + i];
#pragma unroll
              for (int i = 0; i < 3; i++) Trel[i] = cst[C_TREL + i];
              ta = -A.meta[b1].first_child - 1; tb = -B.meta[b2].first_child - 1;
              dTri = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * ta, B.tris + (size_t)TRI_STRIDE * tb, p, qq);
              // the leaf's step bound (only consumed if the leaf lowers the distance; pure, so evaluated here)
              double r1[9], tt1[3], w1[3], w2[3], S1[3], S2[3], tmp[3];
#pragma unroll
              for (int i = 0; i < 9; i++) r1[i] = cst[C_R1 + i];
#pragma unroll
              for (int i = 0; i < 3; i++) tt1[i] = cst[C_TT1 + i];
              m_v(tmp, r1, p); v_add(w1, tmp, tt1);
              m_v(tmp, r1, qq); v_add(w2, tmp, tt1);
              v_sub(S1, w2, w1);
              v_normalize(S1);  // S2 = S1 * -1 normalises to exactly -S1 (see c2a_motion.cuh)
              S2[0] = -S1[0]; S2[1] = -S1[1]; S2[2] = -S1[2];
              Motion m;
#pragma unroll
              for (int i = 0; i < 3; i++) { m.cv[i] = cst[C_CV1 + i]; m.axis[i] = cst[C_AX1 + i]; }
              m.w = cst[C_W1];
              const double mb1 = motion_bound_leaf_unit(m, __ldg(A.geom + (size_t)b1 * GEOM_STRIDE + 15), S1);
#pragma unroll
              for (int i = 0; i < 3; i++) { m.cv[i] = cst[C_CV2 + i]; m.axis[i] = cst[C_AX2 + i]; }
              m.w = cst[C_W2];
              const double mb2 = motion_bound_leaf_unit(m, __ldg(B.geom + (size_t)b2 * GEOM_STRIDE + 15), S2);
              leaf_mt = (dTri) / (mb1 + mb2);
              if (leaf_mt < 0.0) leaf_mt = 0.0;
            }
            __syncwarp();
            // Only a leaf with dTri <= Dw can still lower the distance (Dw never grows), whatever the order the events
            // are resolved in: the others need no slot in the open-leaf list (their triangle test is counted from their
            // record in the fold)
            const unsigned em = __ballot_sync(FULL, go);
            const bool keep = go && !seq && dTri <= Dw;
            const unsigned km = __ballot_sync(FULL, keep);
            const int nfree = 128 - (__popc(ulm0) + __popc(ulm1) + __popc(ulm2) + __popc(ulm3));
            if (!seq && __popc(km) > nfree) { redo = true; break; }
            int slot = 0;
            if (keep)
            {
              int k = __popc(km & ((1u << lane) - 1u));
              const unsigned f0 = ~ulm0, f1 = ~ulm1, f2 = ~ulm2, f3 = ~ulm3;
              const int n0 = __popc(f0), n1 = __popc(f1), n2 = __popc(f2);
              if (k < n0) slot = nth_bit32(f0, k);
              else if (k < n0 + n1) slot = 32 + nth_bit32(f1, k - n0);
              else if (k < n0 + n1 + n2) slot = 64 + nth_bit32(f2, k - n0 - n1);
              else slot = 96 + nth_bit32(f3, k - n0 - n1 - n2);
            }
            if (seq)
            {
              // direct bookkeeping: the single leaf (lane 0) is applied at once
              const double dT = __shfl_sync(FULL, dTri, 0), lm = __shfl_sync(FULL, leaf_mt, 0);
              const int ta0 = __shfl_sync(FULL, ta, 0), tb0 = __shfl_sync(FULL, tb, 0);
              step_ntri++;
              if (dT <= Dw)
              {
                Dw = dT;
                if (lm <= step_mint) step_mint = lm;
                step_lastA = ta0; step_lastB = tb0;
                have_best = true;
#pragma unroll
                for (int i = 0; i < 3; i++) { best_p[i] = __shfl_sync(FULL, p[i], 0); best_q[i] = __shfl_sync(FULL, qq[i], 0); }
              }
            }
            else if (keep)
            {
              ul_key[slot] = key; ul_mpar[slot] = mpar; ul_val[slot] = val; ul_dtri[slot] = dTri;
              double2 *lo = reinterpret_cast<double2 *>(leafout + (size_t)slot * WIDE_LEAFOUT_DOUBLES);
              lo[0] = make_double2(p[0], p[1]); lo[1] = make_double2(p[2], qq[0]);
              lo[2] = make_double2(qq[1], qq[2]); lo[3] = make_double2(leaf_mt, __hiloint2double(ta, tb));
            }
            if (!seq)
            {
              // mark the slots taken (uniform)
              const unsigned s0 = __ballot_sync(FULL, keep && slot < 32), s1 = __ballot_sync(FULL, keep && slot >= 32 && slot < 64);
              const unsigned s2 = __ballot_sync(FULL, keep && slot >= 64 && slot < 96), s3 = __ballot_sync(FULL, keep && slot >= 96);
              // (ballots give lanes, not slots: rebuild the slot masks by OR-reduction)
              unsigned m0 = (keep && slot < 32) ? (1u << slot) : 0u, m1 = (keep && slot >= 32 && slot < 64) ? (1u << (slot - 32)) : 0u;
              unsigned m2 = (keep && slot >= 64 && slot < 96) ? (1u << (slot - 64)) : 0u, m3 = (keep && slot >= 96) ? (1u << (slot - 96)) : 0u;
              if (s0) m0 = __reduce_or_sync(FULL, m0); else m0 = 0;
              if (s1) m1 = __reduce_or_sync(FULL, m1); else m1 = 0;
              if (s2) m2 = __reduce_or_sync(FULL, m2); else m2 = 0;
              if (s3) m3 = __reduce_or_sync(FULL, m3); else m3 = 0;
              ulm0 |= m0; ulm1 |= m1; ulm2 |= m2; ulm3 |= m3;
            }
            // the rest of the waiting list moves down
            const int rest = npl - n_take;
            unsigned long long k0 = 0, k1 = 0; double a0 = 0, a1 = 0, v0 = 0, v1 = 0; int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
            if (lane < rest) { k0 = pl_key[n_take + lane]; a0 = pl_mpar[n_take + lane]; v0 = pl_val[n_take + lane]; x0 = pl_b1[n_take + lane]; y0 = pl_b2[n_take + lane]; }
            if (lane + 32 < rest) { k1 = pl_key[n_take + lane + 32]; a1 = pl_mpar[n_take + lane + 32]; v1 = pl_val[n_take + lane + 32]; x1 = pl_b1[n_take + lane + 32]; y1 = pl_b2[n_take + lane + 32]; }
            __syncwarp();
            if (lane < rest) { pl_key[lane] = k0; pl_mpar[lane] = a0; pl_val[lane] = v0; pl_b1[lane] = x0; pl_b2[lane] = y0; }
            if (lane + 32 < rest) { pl_key[lane + 32] = k1; pl_mpar[lane + 32] = a1; pl_val[lane + 32] = v1; pl_b1[lane + 32] = x1; pl_b2[lane + 32] = y1; }
            npl = rest;
            __syncwarp();
            if ((C2A_WIDE_STATS && args.stats) && lane == 0)
            {
              atomicAdd(args.stats + WS_LEAF_PASSES, 1ull); atomicAdd(args.stats + WS_LEAVES, (unsigned long long)__popc(em));
              atomicAdd(args.stats + WS_CYC_LEAF, (unsigned long long)(clock64() - t_pass));
            }
          }
          else
          {
            // ------------------------------------------------------------ EXPAND pass (C2A.cpp:1192-1276)
            // window: the first Wn entries from the top that still pass under Dw (stale ones are dropped: their
            // step bounds are folded from their records)
            long long ts_a = 0, ts_b = 0, ts_c = 0, ts_d = 0, ts_e = 0, ts_s = 0, ts_r = 0, ts_f = 0;
            (void)ts_a; (void)ts_b; (void)ts_c; (void)ts_d; (void)ts_e; (void)ts_s; (void)ts_r; (void)ts_f;
            const int idx = sp - 1 - lane;
            double Mv = INF;
            if (idx >= 0) Mv = stk[(size_t)idx * ENTRY_DOUBLES + 12];
            bool ok;
            if (seq)
            {
              // one pair per round: the descend test proper, evaluated when the pair is popped (C2A.cpp:1284-1345)
              ok = false;
              if (lane == 0)
              {
                const double mt_e = stk[(size_t)idx * ENTRY_DOUBLES + 15];
                const bool is_root = Mv == -INF;
                ok = is_root || (mt_e < upb && ((Mv < (Dw - abs_err)) || (Mv * (1 + rel_err) < Dw)));
                if (!ok && mt_e < step_mint) step_mint = mt_e;
              }
              step_mint = __shfl_sync(FULL, step_mint, 0);
            }
            else ok = idx >= 0 && Mv < Dw;
            unsigned mask = __ballot_sync(FULL, ok);
            int consumed = seq ? 1 : (sp < 32 ? sp : 32);
            if (!seq && __popc(mask) > Wn)
            {
              const int pw = nth_bit32(mask, Wn - 1);
              consumed = pw + 1;
              mask &= (pw == 31) ? FULL : ((2u << pw) - 1u);
            }
            const int sp_old = sp;
            sp -= consumed;
            const int npairs = __popc(mask);
            const int pair = lane >> 1, c = lane & 1;
            const bool have = pair < npairs;
            // arenas: room for this round's records and pushes?
            if (!seq && (nrec + 2 * npairs > args.rec_cap)) { redo = true; break; }
            if (sp + 2 * npairs > args.stack_cap) { redo = true; seq = true; break; }  // (cannot happen with the caps the host chooses)

            WIDE_TS(ts_a, npairs);
            double R[9], T[3], Mpar = 0, d = 0, mt = 0, val = INF;
            unsigned long long key = 0, ckey = 0;
            int n1 = 0, n2 = 0;
            bool my_leafpair = false, popped_leaf = false;
            NodeMeta cm1, cm2;
            cm1.first_child = 0; cm2.first_child = 0; cm1.size = 0; cm2.size = 0;
            if (have)
            {
              const double *e = stk + (size_t)(sp_old - 1 - nth_bit32(mask, pair)) * ENTRY_DOUBLES;
#pragma unroll
              for (int i = 0; i < 6; i++)
              {
                const double2 v = *reinterpret_cast<const double2 *>(e + 2 * i);
                if (2 * i < 9) R[2 * i] = v.x; else T[2 * i - 9] = v.x;
                if (2 * i + 1 < 9) R[2 * i + 1] = v.y; else T[2 * i + 1 - 9] = v.y;
              }
              const double2 mi = *reinterpret_cast<const double2 *>(e + 12);
              Mpar = mi.x;  // the popped pair's M (seq: its raw d, unused below)
              n1 = __double2hiint(mi.y); n2 = __double2loint(mi.y);
              key = (unsigned long long)__double_as_longlong(e[14]);
              cm1 = load_meta(A.meta + n1); cm2 = load_meta(B.meta + n2);
              const bool l1 = cm1.first_child < 0, l2 = cm2.first_child < 0;
              popped_leaf = l1 && l2;  // (one pair per round only: wide rounds keep leaf pairs off the stack)
            }
            WIDE_TS(ts_b, cm1.size + cm2.size + R[0]);
            ts_c = ts_d = ts_e = ts_b;
            if (have && !popped_leaf)
            {
              const bool l1 = cm1.first_child < 0, l2 = cm2.first_child < 0;
              const double *gs, *gt, *rl;
              double Rc[9], Tc[3];
              if (l2 || (!l1 && (cm1.size > cm2.size)))
              {
                // expansion of side 1, C2A.cpp:1194-1209
                n1 = cm1.first_child + c;
                cm1 = load_meta(A.meta + n1);
                gs = A.geom + (size_t)n1 * GEOM_STRIDE; gt = B.geom + (size_t)n2 * GEOM_STRIDE;
                rl = A.rloc + (size_t)n1 * RLOC_STRIDE;
                double Rn[9], Tn[3], Tt[3];
                load_node_rt(Rn, Tn, gs);
                mt_m(Rc, Rn, R); v_sub(Tt, T, Tn); mt_v(Tc, Rn, Tt);
              }
              else
              {
                // expansion of side 2, C2A.cpp:1211-1225
                n2 = cm2.first_child + c;
                cm2 = load_meta(B.meta + n2);
                gs = A.geom + (size_t)n1 * GEOM_STRIDE; gt = B.geom + (size_t)n2 * GEOM_STRIDE;
                rl = A.rloc + (size_t)n1 * RLOC_STRIDE;
                double Rn[9], Tn[3];
                load_node_rt(Rn, Tn, gt);
                m_m(Rc, R, Rn); m_v_p(Tc, R, Tn, T);
              }
#pragma unroll
              for (int i = 0; i < 9; i++) R[i] = Rc[i];
              T[0] = Tc[0]; T[1] = Tc[1]; T[2] = Tc[2];
              WIDE_TS(ts_c, R[0] + R[4] + R[8] + T[0] + T[1] + T[2]);
              // child BV test (C2A.cpp:1237-1276)
              prefetch_l1(rl); prefetch_l1(rl + 8);
              double S[3];
              const double2 la = __ldg(reinterpret_cast<const double2 *>(gs + 12)), ra2 = __ldg(reinterpret_cast<const double2 *>(gs + 14));
              const double2 lb = __ldg(reinterpret_cast<const double2 *>(gt + 12)), rb2 = __ldg(reinterpret_cast<const double2 *>(gt + 14));
              d = rss_rect_dist(R, T, la.x, la.y, lb.x, lb.y, S);
              d -= (ra2.x + rb2.x);
              d = (d < 0.0) ? 0.0 : d;
              WIDE_TS(ts_d, d);
              if (d != 0.0)
              {
                double Rl[9], tmp[3], S1[3], S2[3], r1[9];
                load9v(Rl, rl);
                m_v(tmp, Rl, S);
#pragma unroll
                for (int i = 0; i < 9; i++) r1[i] = cst[C_R1 + i];
                m_v(S1, r1, tmp);
                v_normalize(S1);
                S2[0] = -S1[0]; S2[1] = -S1[1]; S2[2] = -S1[2];
                Motion m;
#pragma unroll
                for (int i = 0; i < 3; i++) { m.cv[i] = cst[C_CV1 + i]; m.axis[i] = cst[C_AX1 + i]; }
                m.w = cst[C_W1];
                const double mb1 = motion_bound_bv_unit(m, ra2.y, S1);
#pragma unroll
                for (int i = 0; i < 3; i++) { m.cv[i] = cst[C_CV2 + i]; m.axis[i] = cst[C_AX2 + i]; }
                m.w = cst[C_W2];
                const double mb2 = motion_bound_bv_unit(m, rb2.y, S2);
                mt = (d) / (mb1 + mb2);
                if (mt <= 0) mt = 0.0;
              }
              my_leafpair = cm1.first_child < 0 && cm2.first_child < 0;
              val = (mt < upb) ? d : INF;
              WIDE_TS(ts_e, mt);
            }
            __syncwarp();
            // the pair exchanges its two tests; the closer child is visited first, ties visit 'a' first (d2 < d1)
            const double d_o = __shfl_xor_sync(FULL, d, 1);
            WIDE_TS(ts_s, d_o);
            ts_r = ts_f = ts_s;
            const bool c_first = c ? (d < d_o) : (d_o < d);
            const int j = (c == 1) == c_first ? 0 : 1;   // 0: visited first
            if (seq)
            {
              // direct bookkeeping (lane pair 0): C2A.cpp:1279-1351 with the far child re-tested when popped
              if (npairs > 0 && __shfl_sync(FULL, popped_leaf ? 1 : 0, 0))
              {
                // a leaf pair: the LEAF pass of the next round tests it
                if (lane == 0) { pl_key[0] = 0; pl_mpar[0] = 0; pl_val[0] = 0; pl_b1[0] = n1; pl_b2[0] = n2; }
                npl = 1;
                __syncwarp();
              }
              else if (npairs > 0)
              {
                step_nbv += 2;
                const bool pass = have && mt < upb && ((d < (Dw - abs_err)) || (d * (1 + rel_err) < Dw));
                double fm = (have && !pass) ? mt : INF;
                const double fm_o = __shfl_xor_sync(FULL, fm, 1);
                fm = fm_o < fm ? fm_o : fm;
                fm = __shfl_sync(FULL, fm, 0);
                if (fm < step_mint) step_mint = fm;
                const unsigned pm = __ballot_sync(FULL, pass) & 3u;
                const bool partner_pushes = (pm >> (c ^ 1)) & 1u;
                const int total = __popc(pm);
                if (pass)
                {
                  const int rank = (j == 1 && partner_pushes) ? 1 : 0;
                  double *e = stk + (size_t)(sp + total - 1 - rank) * ENTRY_DOUBLES;
                  double2 *e2 = reinterpret_cast<double2 *>(e);
                  e2[0] = make_double2(R[0], R[1]); e2[1] = make_double2(R[2], R[3]);
                  e2[2] = make_double2(R[4], R[5]); e2[3] = make_double2(R[6], R[7]);
                  e2[4] = make_double2(R[8], T[0]); e2[5] = make_double2(T[1], T[2]);
                  e2[6] = make_double2(d, __hiloint2double(n1, n2)); e2[7] = make_double2(0.0, mt);
                }
                sp += total;
                __syncwarp();
              }
            }
            else
            {
              const int depth = (int)(key & 0x3full);
              if (__any_sync(FULL, have && depth >= WIDE_KEY_LEVELS)) { redo = true; break; }
              if (lane == 0 && nrnd < WIDE_RND) rnd_tab[nrnd] = make_int2(nrec, nev);
              nrnd++;
              if (have)
              {
                ckey = (key & ~0xffull) | ((unsigned long long)j << (63 - depth)) | (unsigned long long)(depth + 1) | (my_leafpair ? 0x80ull : 0ull);
                // (streaming stores: a record is written once and read once, by the fold; 5 GB of record arenas must not
                // push the stacks and the models out of L2)
                double2 *rr = reinterpret_cast<double2 *>(recs + (size_t)(nrec + lane) * 4);
                __stcs(rr, make_double2(__longlong_as_double((long long)ckey), Mpar));
                __stcs(rr + 1, make_double2(val, mt));
              }
              nrec += 2 * npairs;
              WIDE_TS(ts_r, nrec);
              const bool pass = have && val < Dw;   // (the popped pair's M < Dw already)
              const double Mc = Mpar > val ? Mpar : val;
              // waiting leaves
              const unsigned lm = __ballot_sync(FULL, pass && my_leafpair);
              if (pass && my_leafpair)
              {
                const int s = npl + __popc(lm & ((1u << lane) - 1u));
                pl_key[s] = ckey; pl_mpar[s] = Mpar; pl_val[s] = val; pl_b1[s] = n1; pl_b2[s] = n2;
              }
              npl += __popc(lm);
              // stack pushes in visiting order: rank 0 (the window's first pair's first child) ends up on top
              const unsigned pm = __ballot_sync(FULL, pass && !my_leafpair);
              const int total = __popc(pm);
              const bool partner_pushes = (pm >> (lane ^ 1)) & 1u;
              const int rank = __popc(pm & ((1u << (2 * pair)) - 1u)) + ((j == 1 && partner_pushes) ? 1 : 0);
              if (pass && !my_leafpair)
              {
                double *e = stk + (size_t)(sp + total - 1 - rank) * ENTRY_DOUBLES;
                double2 *e2 = reinterpret_cast<double2 *>(e);
                e2[0] = make_double2(R[0], R[1]); e2[1] = make_double2(R[2], R[3]);
                e2[2] = make_double2(R[4], R[5]); e2[3] = make_double2(R[6], R[7]);
                e2[4] = make_double2(R[8], T[0]); e2[5] = make_double2(T[1], T[2]);
                e2[6] = make_double2(Mc, __hiloint2double(n1, n2)); e2[7] = make_double2(__longlong_as_double((long long)ckey), mt);
              }
              sp += total;
              if (total > 0)
              {
                const unsigned tm = __ballot_sync(FULL, pass && !my_leafpair && rank == 0);
                top_key = shfl_u64(FULL, ckey, __ffs(tm) - 1);
                top_known = true;
              }
              __syncwarp();
              WIDE_TS(ts_f, sp);
            }
            if ((C2A_WIDE_STATS && args.stats) && lane == 0)
            {
              atomicAdd(args.stats + WS_X_SYNC, (unsigned long long)(ts_s - ts_e)); atomicAdd(args.stats + WS_X_RECORD, (unsigned long long)(ts_r - ts_s));
              atomicAdd(args.stats + WS_X_PUSH, (unsigned long long)(ts_f - ts_r));
              atomicAdd(args.stats + WS_ROUNDS, 1ull); atomicAdd(args.stats + WS_TESTS, (unsigned long long)(2 * npairs));
              atomicAdd(args.stats + WS_CYC_EXPAND, (unsigned long long)(clock64() - t_pass));
              atomicAdd(args.stats + WS_X_SELECT, (unsigned long long)(ts_a - t_pass)); atomicAdd(args.stats + WS_X_POP, (unsigned long long)(ts_b - ts_a));
              atomicAdd(args.stats + WS_X_XFORM, (unsigned long long)(ts_c - ts_b)); atomicAdd(args.stats + WS_X_RECT, (unsigned long long)(ts_d - ts_c));
              atomicAdd(args.stats + WS_X_BOUND, (unsigned long long)(ts_e - ts_d));
            }
          }
          if (seq) continue;

          // ---------------------------------------------------------------- resolve events
          // everything that precedes F (the top of the stack, the waiting leaf pairs) has been evaluated
          const long long t_res = (C2A_WIDE_STATS && args.stats) ? clock64() : 0;
          if (ulm0 | ulm1 | ulm2 | ulm3)
          {
            unsigned long long F = ~0ull;
            if (sp > 0) F = top_known ? top_key : (unsigned long long)__double_as_longlong(__ldcg(stk + (size_t)(sp - 1) * ENTRY_DOUBLES + 14));
            if (npl > 0)
            {
              unsigned long long k = ~0ull;
              if (lane < npl) k = pl_key[lane];
              if (lane + 32 < npl) { const unsigned long long k2 = pl_key[lane + 32]; k = k2 < k ? k2 : k; }
              k = warp_min_u64(k);
              F = k < F ? k : F;
            }
            // my open leaves (slots lane, lane + 32, ...): those that can no longer lower the distance under the current Dw
            // are dropped whatever their position (Dw never grows); of the others, the ones below F are candidates
            unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull, k3 = ~0ull;
            bool dead0 = false, dead1 = false, dead2 = false, dead3 = false;
            if ((ulm0 >> lane) & 1u) { const double m_ = ul_mpar[lane], v_ = ul_val[lane]; dead0 = !((m_ > v_ ? m_ : v_) < Dw && ul_dtri[lane] <= Dw); if (!dead0) { const unsigned long long k = ul_key[lane]; if (k < F) k0 = k; } }
            if ((ulm1 >> lane) & 1u) { const double m_ = ul_mpar[32 + lane], v_ = ul_val[32 + lane]; dead1 = !((m_ > v_ ? m_ : v_) < Dw && ul_dtri[32 + lane] <= Dw); if (!dead1) { const unsigned long long k = ul_key[32 + lane]; if (k < F) k1 = k; } }
            if ((ulm2 >> lane) & 1u) { const double m_ = ul_mpar[64 + lane], v_ = ul_val[64 + lane]; dead2 = !((m_ > v_ ? m_ : v_) < Dw && ul_dtri[64 + lane] <= Dw); if (!dead2) { const unsigned long long k = ul_key[64 + lane]; if (k < F) k2 = k; } }
            if ((ulm3 >> lane) & 1u) { const double m_ = ul_mpar[96 + lane], v_ = ul_val[96 + lane]; dead3 = !((m_ > v_ ? m_ : v_) < Dw && ul_dtri[96 + lane] <= Dw); if (!dead3) { const unsigned long long k = ul_key[96 + lane]; if (k < F) k3 = k; } }
            ulm0 &= ~__ballot_sync(FULL, dead0); ulm1 &= ~__ballot_sync(FULL, dead1);
            ulm2 &= ~__ballot_sync(FULL, dead2); ulm3 &= ~__ballot_sync(FULL, dead3);
            while (true)
            {
              // the smallest key among them: two hardware min-reductions (high, then low word)
              unsigned long long mk = k0; int ms = lane;
              if (k1 < mk) { mk = k1; ms = 32 + lane; }
              if (k2 < mk) { mk = k2; ms = 64 + lane; }
              if (k3 < mk) { mk = k3; ms = 96 + lane; }
              const unsigned hi = (unsigned)(mk >> 32), lo = (unsigned)mk;
              const unsigned mhi = __reduce_min_sync(FULL, hi);
              const unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
              if (mhi == 0xffffffffu && mlo == 0xffffffffu) break;   // (no key is all ones: the depth byte is < 64)
              const unsigned owner = __ballot_sync(FULL, hi == mhi && lo == mlo);
              const int s = __shfl_sync(FULL, ms, __ffs(owner) - 1);
              const unsigned long long best = ((unsigned long long)mhi << 32) | mlo;
              const double mpar = ul_mpar[s], val = ul_val[s], dTri = ul_dtri[s];
              const double M = mpar > val ? mpar : val;
              if (M < Dw && dTri <= Dw)
              {
                if (dTri != 0.0 && !(mpar < dTri)) { redo = true; break; }  // ancestor anomaly: redo the step sequentially
                if (nev >= WIDE_EV) { redo = true; break; }
                Dw = dTri;
                if (lane == 0) { ev_key[nev] = best; ev_d[nev] = dTri; }
                nev++;
                const double2 *lo2 = reinterpret_cast<const double2 *>(leafout + (size_t)s * WIDE_LEAFOUT_DOUBLES);
                const double2 l0 = __ldcg(lo2), l1 = __ldcg(lo2 + 1), l2 = __ldcg(lo2 + 2), l3 = __ldcg(lo2 + 3);
                best_p[0] = l0.x; best_p[1] = l0.y; best_p[2] = l1.x; best_q[0] = l1.y; best_q[1] = l2.x; best_q[2] = l2.y;
                if (l3.x <= step_mint) step_mint = l3.x;
                step_lastA = __double2hiint(l3.y); step_lastB = __double2loint(l3.y);
                have_best = true;
              }
              const unsigned bit = 1u << (s & 31);
              if (s < 32) ulm0 &= ~bit; else if (s < 64) ulm1 &= ~bit; else if (s < 96) ulm2 &= ~bit; else ulm3 &= ~bit;
              if ((s & 31) == lane) { if (s < 32) k0 = ~0ull; else if (s < 64) k1 = ~0ull; else if (s < 96) k2 = ~0ull; else k3 = ~0ull; }
            }
          }
          __syncwarp();
          if ((C2A_WIDE_STATS && args.stats) && lane == 0) atomicAdd(args.stats + WS_CYC_RESOLVE, (unsigned long long)(clock64() - t_res));
          if (redo) break;
        }
        if (redo)
        {
          if ((C2A_WIDE_STATS && args.stats) && lane == 0) atomicAdd(args.stats + WS_REDO, 1ull);
          seq = true;
          __syncwarp();
          // the root entry was consumed: rewrite its M slot is not needed (entry 0 is intact: pops do not erase), but
          // pushes may have overwritten it -- rebuild it from the constants' source: redo the set-up cheaply
          if (lane == 0)
          {
            double g1[12], g2[12], Rt[9], R[9], T[3], Tt[3], Rrel[9], Trel[3];
#pragma unroll
            for (int i = 0; i < 9; i++) Rrel[i] = cst[C_RREL + i];
#pragma unroll
            for (int i = 0; i < 3; i++) Trel[i] = cst[C_TREL + i];
#pragma unroll
            for (int i = 0; i < 12; i++) { g1[i] = __ldg(A.geom + i); g2[i] = __ldg(B.geom + i); }
            m_m(Rt, Rrel, g2);
            mt_m(R, g1, Rt);
            m_v_p(Tt, Rrel, &g2[9], Trel);
            v_sub(Tt, Tt, &g1[9]);
            mt_v(T, g1, Tt);
            double2 *e2 = reinterpret_cast<double2 *>(stk);
            e2[0] = make_double2(R[0], R[1]); e2[1] = make_double2(R[2], R[3]);
            e2[2] = make_double2(R[4], R[5]); e2[3] = make_double2(R[6], R[7]);
            e2[4] = make_double2(R[8], T[0]); e2[5] = make_double2(T[1], T[2]);
            e2[6] = make_double2(-INF, __hiloint2double(0, 0)); e2[7] = make_double2(__longlong_as_double(0ll), 0.0);
          }
          __syncwarp();
          continue;
        }
        if (seq)
        {
          step_dist = Dw;
          if ((C2A_WIDE_STATS && args.stats) && lane == 0) atomicAdd(args.stats + WS_SEQ_STEPS, 1ull);
          break;
        }

        // ------------------------------------------------------------------ FOLD
        const long long t_fold = (C2A_WIDE_STATS && args.stats) ? clock64() : 0;
        {
          int f_nbv = 0, f_ntri = root_leaf ? 1 : 0;
          double f_mint = step_mint;
          // records were written round by round; the events resolved before a record's round precede its node, so the
          // search for "the last event before this position" starts there and is short
          const int nr = nrnd < WIDE_RND ? nrnd : WIDE_RND;
          int rc = 0;
          double2 n0 = make_double2(0, 0), n1 = make_double2(0, 0);
          if (lane < nrec) { n0 = __ldcs(reinterpret_cast<const double2 *>(recs + (size_t)lane * 4)); n1 = __ldcs(reinterpret_cast<const double2 *>(recs + (size_t)lane * 4) + 1); }
          for (int i = lane; i < nrec; i += 32)
          {
            const double2 r0 = n0, r1 = n1;
            if (i + 32 < nrec) { n0 = __ldcs(reinterpret_cast<const double2 *>(recs + (size_t)(i + 32) * 4)); n1 = __ldcs(reinterpret_cast<const double2 *>(recs + (size_t)(i + 32) * 4) + 1); }
            while (rc + 1 < nr && rnd_tab[rc + 1].x <= i) rc++;
            const unsigned long long key = (unsigned long long)__double_as_longlong(r0.x);
            const int depth = (int)(key & 0x3full);
            const unsigned long long pkey = ((key & ~0xffull) & ~(1ull << (64 - depth))) | (unsigned long long)(depth - 1);
            // D in force at the parent's and at the node's own position: the last event before it
            int lo = rnd_tab[rc].y;
            while (lo < nev && ev_key[lo] < pkey) lo++;
            const double Dp = lo ? ev_d[lo - 1] : seed_dist;
            if (!(r0.y < Dp)) continue;  // the parent was not visited
            f_nbv++;
            while (lo < nev && ev_key[lo] < key) lo++;   // (own position is not before the parent's)
            const double Dn = lo ? ev_d[lo - 1] : seed_dist;
            if (!(r1.x < Dn)) { if (r1.y < f_mint) f_mint = r1.y; }
            else if (key & 0x80ull) f_ntri++;
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1)
          {
            f_nbv += __shfl_xor_sync(FULL, f_nbv, o);
            f_ntri += __shfl_xor_sync(FULL, f_ntri, o);
            const double m2 = __shfl_xor_sync(FULL, f_mint, o);
            f_mint = m2 < f_mint ? m2 : f_mint;
          }
          f_ntri -= root_leaf ? 31 : 0;  // (the root leaf was counted by every lane)
          step_nbv = f_nbv; step_ntri = f_ntri; step_mint = f_mint; step_dist = Dw;
        }
        if ((C2A_WIDE_STATS && args.stats) && lane == 0)
        {
          atomicAdd(args.stats + WS_EVENTS, (unsigned long long)nev);
          atomicAdd(args.stats + WS_CYC_FOLD, (unsigned long long)(clock64() - t_fold));
        }
        break;
      }

      // ---- the step's results
      dist = step_dist; mint = step_mint; nbv += step_nbv; ntri += step_ntri; lastA = step_lastA; lastB = step_lastB;
      if (have_best && args.out.p1p2 && lane == 0)
      {
#pragma unroll
        for (int i = 0; i < 3; i++) { args.out.p1p2[6 * q + i] = best_p[i]; args.out.p1p2[6 * q + 3 + i] = best_q[i]; }
      }

      // ---- C2A_QueryTimeOfContact's loop, C2A.cpp:2053-2123 (uniform across the warp)
      bool finished = false, hit = false;
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
            else { lastLamda = lamda; numCA++; upb = 1.0 - lamda; }
          }
        }
      }
      if (finished)
      {
        // C2A.cpp:2125-2143 and the pose outputs of C2A_Solve :2411-2429
        if (lane == 0)
        {
          double toc = 0.0;
          const c2a_b200_results &o = args.out;
          if (hit)
          {
            toc = lastLamda;
            if (toc >= 1 - args.tol_t) toc = 0;
            if (o.pose_toc)
            {
              double R[9], T[3];
              motion_pose_nl(rec, toc, R, T);
#pragma unroll
              for (int i = 0; i < 9; i++) o.pose_toc[24 * q + i] = R[i];
#pragma unroll
              for (int i = 0; i < 3; i++) o.pose_toc[24 * q + 9 + i] = T[i];
              motion_pose_nl(rec + MOTION_DOUBLES, toc, R, T);
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
          if (o.last_tri) { o.last_tri[2 * q] = lastA; o.last_tri[2 * q + 1] = lastB; }
          if (args.trace) args.trace[2 * q + 1] = global_ns();
        }
        break;
      }
    }
  }
  if ((C2A_WIDE_STATS && args.stats) && threadIdx.x == 0) atomicMax(args.stats + WS_T_LAST, global_ns());
}

}  // namespace c2a
