// Persistent warp-cooperative kernel: batched controlled conservative advancement (sm_100a, FP64).
//
// What runs on the device, per query (reference paths relative to /root/reference):
//   C2A_QueryTimeOfContact  (the CA loop)        C2A/src/C2A.cpp:1987-2146
//   C2A_TimeOfContactStep   (per-step set-up)    C2A/src/C2A.cpp:1778-1931
//   TOCStepRecurse_Dis      (BVTT traversal)     C2A/src/C2A.cpp:1114-1354
//   pose outputs of C2A_Solve                    C2A/src/C2A.cpp:2411-2429
// with no host round trip between CA iterations.
//
// Execution model.  The traversal result is order dependent (res->distance shrinks as leaves are
// visited and gates pruning, C2A.cpp:1281-1351), so every query commits node pairs in the reference's
// depth-first order and parallelism is taken ACROSS queries and across the two child tests of one
// expansion.  A warp owns a pool of Q query slots whose state lives in shared memory; lanes are
// anonymous workers.  Each trip round the main loop the warp ballots the slot states and runs ONE
// phase with as many lanes as it can fill:
//   EXPAND   16 slots at a time, a lane PAIR per slot: both lanes fetch the slot's next node pair
//            (the current entry in shared memory, else pop the slot's stack until an entry passes
//            the descend test), then lane c runs child test c (transform, RSS distance, motion
//            bounds); the pair exchanges (d, mint) by shuffle and commits: near child -> current
//            entry, far child -> stack (only if it passes the descend test now: the distance only
//            shrinks, so a child that fails now fails later and its step bound is folded at once).
//   LEAF     one lane per slot waiting on a triangle pair: triangle distance + leaf motion bounds.
//   ADVANCE  one lane per slot whose step ended or that is empty: CA-loop bookkeeping, result
//            write-out, claim of the next query from a global atomic counter, next step's set-up.
// Traversal stacks live in global memory (L2-resident), one 128-byte entry per pending far child.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/c2a_b200.h"
#include "c2a_geom.cuh"
#include "c2a_motion.cuh"

namespace c2a {

// ---- device-resident model ---------------------------------------------------------------
// geom  [n][16]  R(9) Tr(3) l(2) r ang_radius      one 128-byte line per node; children adjacent
// rloc  [n][10]  R_loc(9) pad (only read when a BV distance is non-zero)
// meta  [n]      {GetSize() = sqrt(l0^2+l1^2)+2r precomputed (PQP BV::GetSize), first_child}
// tris  [n][10]  p1 p2 p3 pad
constexpr int GEOM_STRIDE = 16;
constexpr int RLOC_STRIDE = 10;   // R_loc(9) + pad: 80-byte records, so that they can be fetched with 128-bit loads
constexpr int TRI_STRIDE = 10;    // p1 p2 p3 (9) + pad, likewise
constexpr int OBB_STRIDE = 6;    // OBB half-dimensions d(3) + centre To(3) (C2A_Collide only)
struct NodeMeta { double size; int first_child; int pad; };
struct DevModel
{
  const double *geom;
  const double *rloc;
  const NodeMeta *meta;
  const double *tris;
  int n_nodes, n_tris;
};

constexpr int ENTRY_DOUBLES = 16;  // R(9) T(3) d mint {b1,b2} pad  -> 128 B
#ifndef C2A_WPB
#define C2A_WPB 4
#endif
#ifndef C2A_MINB
#define C2A_MINB 2
#endif
constexpr int WARPS_PER_BLOCK = C2A_WPB;
constexpr int BLOCK_THREADS = 32 * WARPS_PER_BLOCK;
#ifndef C2A_Q
#define C2A_Q 48
#endif
constexpr int Q = C2A_Q;              // query slots per warp: more than one EXPAND pass (16) plus one LEAF pass can use,
                                   // so that leaves pile up to a full 32-lane LEAF pass while expansion stays fed

struct BatchArgs
{
  DevModel A, B;
  const double *motions;  // [n][2][MOTION_DOUBLES], see c2a_motion.cuh
  const int *seedA, *seedB;
  long long n;
  double tol_d, tol_t;
  c2a_b200_results out;
  unsigned long long *counter;
  double *stacks;         // [n_slots][stack_entries][ENTRY_DOUBLES]
  int stack_entries;
  const int *order;       // optional [n]: the k-th claim takes query order[k] (longest-expected first); NULL = identity
  const double *step_in;  // optional [n][STEP_IN_DOUBLES]: single-step mode (C2A_TimeOfContactStep), see below
  unsigned long long *stats;  // optional [14], see c2a_b200_phase_stats; NULL = off
  unsigned long long *trace;  // optional [n][2]: globaltimer at claim / at result write-out (development aid)
  // Hand-over to c2a_wide_kernel (c2a_wide.cuh; spill_recs = NULL: off).  Once the claim queue is empty and a warp holds
  // at most spill_live queries, a query past its fifth CA step is written out at its next step boundary (where its
  // state is 64 bytes: no stack) and the slot retires; the wide kernel, launched next on the stream, finishes it with
  // parallelism inside the traversal.
  double *spill_recs;                 // [SPILL_BUCKETS][spill_cap][MB_DOUBLES]
  unsigned long long *spill_count;    // control words, SPILL_CTL_*: [0..2] records reserved per list, [4..6] records claimed by
                                      // the wide kernel, [8] blocks of this kernel that have started, [9] warps that are done
  long long spill_cap;
  int spill_live;
  int max_slots;                      // query slots a warp uses (<= Q): small batches spread over all warps
};
constexpr int SPILL_BUCKETS = 3, SPILL_CA_HEAVY = 40, SPILL_CA_MID = 16;
constexpr int SPILL_CTL_CLAIMED = 4, SPILL_CTL_STARTED = 8, SPILL_CTL_DONE = 9, SPILL_CTL_WORDS = 32;
constexpr int MB_DOUBLES = 8;  // q, lamda, lastLamda, mint, UpboundTOC, {numCA, nItrs}, {nbv, ntri}, {lastA, lastB}

// Single-step mode (the device side of C2A_TimeOfContactStep, C2A.cpp:1778-1931): per query the caller
// supplies the current poses and the CA-loop state the step reads -- R1(9) T1(3) R2(9) T2(3) numCA
// res->mint(previous step) res->UpboundTOC pad -- and one traversal is run; distance, mint, p1p2 and
// the two counters are written, nothing else.
constexpr int STEP_IN_DOUBLES = 28;

enum SlotState { ST_ADVANCE = 0, ST_TRAVERSE = 1, ST_LEAF = 2, ST_EXIT = 3 };

// per-slot state in shared memory, structure of arrays [field][slot]
enum
{
  F_R1 = 0,      // 9  current rotation of object 1 (objmotion1->transform)
  F_TT1 = 9,     // 3  current translation of object 1
  F_CV1 = 12, F_AX1 = 15, F_W1 = 18,   // motion 1: cv, m_axis, m_angVel
  F_CV2 = 19, F_AX2 = 22, F_W2 = 25,   // motion 2
  F_RREL = 26,   // 9  res->R
  F_TREL = 35,   // 3  res->T
  F_DIST = 38, F_MINT = 39, F_ABS = 40, F_REL = 41, F_UPB = 42,
  F_CUR = 43,    // 12 current entry: R(9) T(3) of the node pair to visit next
  F_LAMDA = 55, F_LASTL = 56,
  F_CURSZ1 = 57, F_CURSZ2 = 58,  // GetSize() of the current entry's two nodes (with I_CURFC1/2: their NodeMeta)
  F_NDBL = 59                    // (res->p1/p2 are outputs only: written straight to the result arrays)
};
enum
{
  I_STATE = 0, I_SP, I_NBV, I_NTRI, I_NUMCA, I_NITRS, I_CURB1, I_CURB2, I_LEAFB1, I_LEAFB2, I_SEEDA, I_SEEDB,
  I_QLO, I_QHI, I_PENDING, I_CURFC1, I_CURFC2, I_LASTA, I_LASTB, I_NINT = 20
};
constexpr size_t WARP_SMEM_BYTES = (size_t)F_NDBL * Q * 8 + (size_t)I_NINT * Q * 4;
constexpr size_t BLOCK_SMEM_BYTES = WARP_SMEM_BYTES * WARPS_PER_BLOCK;

C2A_DEV void load9(double d[9], const double *s)
{
#pragma unroll
  for (int i = 0; i < 9; i++) d[i] = __ldg(s + i);
}
C2A_DEV void load3(double d[3], const double *s) { d[0] = __ldg(s); d[1] = __ldg(s + 1); d[2] = __ldg(s + 2); }
// nine doubles of a 16-byte aligned, padded record (R_loc, triangle) with five 128-bit loads
C2A_DEV void load9v(double d[9], const double *s)
{
  const double2 *s2 = reinterpret_cast<const double2 *>(s);
  const double2 v0 = __ldg(s2), v1 = __ldg(s2 + 1), v2 = __ldg(s2 + 2), v3 = __ldg(s2 + 3), v4 = __ldg(s2 + 4);
  d[0] = v0.x; d[1] = v0.y; d[2] = v1.x; d[3] = v1.y; d[4] = v2.x; d[5] = v2.y; d[6] = v3.x; d[7] = v3.y; d[8] = v4.x;
}
// R(9) + Tr(3) of a node record (128-byte aligned) with six 128-bit loads
C2A_DEV void load_node_rt(double R[9], double T[3], const double *g)
{
  const double2 *g2 = reinterpret_cast<const double2 *>(g);
  const double2 v0 = __ldg(g2), v1 = __ldg(g2 + 1), v2 = __ldg(g2 + 2), v3 = __ldg(g2 + 3), v4 = __ldg(g2 + 4), v5 = __ldg(g2 + 5);
  R[0] = v0.x; R[1] = v0.y; R[2] = v1.x; R[3] = v1.y; R[4] = v2.x; R[5] = v2.y; R[6] = v3.x; R[7] = v3.y; R[8] = v4.x;
  T[0] = v4.y; T[1] = v5.x; T[2] = v5.y;
}
// one 128-bit load of a node's {GetSize(), first_child}
C2A_DEV NodeMeta load_meta(const NodeMeta *p)
{
  const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  NodeMeta m;
  m.size = v.x; m.first_child = __double2loint(v.y); m.pad = 0;
  return m;
}
// per-query seed triangle (res->last_triA/B); an index outside the model falls back to triangle 0 (the state after
// EndModel, C2A_PQP.cpp:401) instead of reading out of bounds -- the host entries reject such seeds with ERR_ARG
C2A_DEV int seed_or_zero(const int *seeds, long long q, int n_tris)
{
  if (!seeds) return 0;
  const int s = seeds[q];
  return ((unsigned)s < (unsigned)n_tris) ? s : 0;
}
C2A_DEV unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
C2A_DEV void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// index of the (n+1)-th set bit of a 32-bit word (n < popc(m)): five branch-free halving steps on popc
// (the __fns intrinsic is a software loop and showed up with 4.5 % of the stall samples)
C2A_DEV int nth_bit32(unsigned m, int n)
{
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1)
  {
    const int c = __popc(m & ((1u << w) - 1u));
    const bool up = n >= c;
    n -= up ? c : 0;
    m = up ? (m >> w) : m;
    pos += up ? w : 0;
  }
  return pos;
}
// index of the (n+1)-th set bit of a 64-bit slot mask
C2A_DEV int nth_slot(unsigned long long m, int n)
{
  const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
  const int nlo = __popc(lo);
  const bool in_lo = n < nlo;
  return (in_lo ? 0 : 32) + nth_bit32(in_lo ? lo : hi, in_lo ? n : n - nlo);
}

__device__ __noinline__ double tri_distance_nl(const double R[9], const double T[3], const double *t1,
                                               const double *t2, double p[3], double q[3])
{
  double a[9], b[9];
  load9(a, t1);
  load9(b, t2);
  return tri_distance(R, T, a, b, p, q);
}

// the same for triangles of an uploaded model (TRI_STRIDE records): 128-bit loads
__device__ __noinline__ double tri_distance_v(const double R[9], const double T[3], const double *t1, const double *t2,
                                              double p[3], double q[3])
{
  double a[9], b[9];
  load9v(a, t1);
  load9v(b, t2);
  return tri_distance(R, T, a, b, p, q);
}

// pose integration is only used by the (rare) ADVANCE phase; keeping it out of line keeps the libm
// sin/cos bodies (~600 SASS instructions per call site) out of the hot loop's instruction footprint
__device__ __noinline__ void motion_pose_nl(const double *rec, double t, double R[9], double T[3])
{
  Motion m;
  motion_load(m, rec);
  motion_pose(m, t, R, T);
}

// The phase counters / per-query timeline (c2a_b200_phase_stats, c2a_b200_query_trace: development aids) are compiled out
// of the product build: merely being there (a clock read and a few predicated atomics per pass, two registers, 2 KB of
// code) cost the kernel 6 % (1 M queries: 7.35 s -> 6.90 s).  scripts/build_variant.py <name> -DC2A_SOLVE_STATS=1 makes a
// library with them (scripts/phase_stats.py, wide_stats.py, tail_stats.py, query_trace.py read them).
#ifndef C2A_SOLVE_STATS
#define C2A_SOLVE_STATS 0
#endif
#define SOLVE_STATS_PTR (C2A_SOLVE_STATS ? args.stats : (unsigned long long *)nullptr)
#define SOLVE_TRACE_PTR (C2A_SOLVE_STATS ? args.trace : (unsigned long long *)nullptr)
__global__ void __launch_bounds__(BLOCK_THREADS, C2A_MINB) c2a_solve_kernel(const BatchArgs args)
{
  extern __shared__ double smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *sd = smem + (size_t)warp * (WARP_SMEM_BYTES / 8);
  int *si = reinterpret_cast<int *>(sd + (size_t)F_NDBL * Q);
  const DevModel &A = args.A, &B = args.B;
  double *const stack_base = args.stacks + ((size_t)(blockIdx.x * WARPS_PER_BLOCK + warp) * Q) * args.stack_entries * ENTRY_DOUBLES;
  const size_t stack_stride = (size_t)args.stack_entries * ENTRY_DOUBLES;
  if (args.spill_count && threadIdx.x == 0) atomicAdd(args.spill_count + SPILL_CTL_STARTED, 1ull);

#define SD(f, s) sd[(f) * Q + (s)]
#define SI(f, s) si[(f) * Q + (s)]

  if (SOLVE_STATS_PTR && threadIdx.x == 0) atomicMin(SOLVE_STATS_PTR + 6, global_ns());  // launch start
  for (int sl = lane; sl < Q; sl += 32)
  {
    SI(I_STATE, sl) = sl < args.max_slots ? ST_ADVANCE : ST_EXIT;
    SI(I_QLO, sl) = -1; SI(I_QHI, sl) = -1;
    SI(I_PENDING, sl) = 0;
  }
  __syncwarp();

  unsigned pass_no = 0;
  while (true)
  {
    pass_no++;
    const int st = SI(I_STATE, lane), st2 = (lane + 32 < Q) ? SI(I_STATE, lane + 32) : (int)ST_EXIT;
    const unsigned long long mT = __ballot_sync(FULL, st == ST_TRAVERSE) | ((unsigned long long)__ballot_sync(FULL, st2 == ST_TRAVERSE) << 32);
    const unsigned long long mL = __ballot_sync(FULL, st == ST_LEAF) | ((unsigned long long)__ballot_sync(FULL, st2 == ST_LEAF) << 32);
    const unsigned long long mA = __ballot_sync(FULL, st == ST_ADVANCE) | ((unsigned long long)__ballot_sync(FULL, st2 == ST_ADVANCE) << 32);
    const int n_live = __popcll(mT | mL | mA);
    if (n_live == 0) break;
    const int nT = __popcll(mT), nL = min(__popcll(mL), 32), nA = min(__popcll(mA), 32);  // a pass serves at most 32 slots
    // Phase choice: a full expansion pass (16 slots x 2 lanes) whenever one is available; otherwise
    // first turn waiting slots back into traversable ones (the larger of the LEAF / ADVANCE groups),
    // and only run a partial expansion pass when nothing is waiting.
    int phase;
    if (nL == 32) phase = ST_LEAF;  // cannot get fuller
    else if (nT >= 16) phase = ST_TRAVERSE;
    else if (nL > 0 || nA > 0) phase = (nL >= nA) ? ST_LEAF : ST_ADVANCE;
    else phase = ST_TRAVERSE;

    const bool steady = n_live == args.max_slots && n_live == Q;  // statistics cover warps whose slots are all live (not the tail)
    if (SOLVE_STATS_PTR && lane == 0 && steady)
    {
      const int k = phase == ST_TRAVERSE ? 0 : (phase == ST_LEAF ? 2 : 4);
      atomicAdd(SOLVE_STATS_PTR + k, 1ull);
      atomicAdd(SOLVE_STATS_PTR + k + 1, (unsigned long long)(phase == ST_TRAVERSE ? (nT >= 5 ? 2 * min(nT, 16) : (nT >= 3 ? 6 * nT : (nT == 2 ? 28 : 30))) : (phase == ST_LEAF ? nL : nA)));  // LEAF: slots served (9 lanes each, 3 per round)
    }
    const long long pass_t0 = SOLVE_STATS_PTR ? clock64() : 0;
    if (phase == ST_TRAVERSE)
    {
      // -------------------------------------------------------------------- EXPAND ----------
      // G lanes per slot.  G = 2 (16 slots per pass) is the plain expansion: lane c runs child test c.
      // With few traversable slots (the tail of a launch, or a single-query call) the idle lanes
      // look ahead: G = 8 / 16 / 32 lanes evaluate the child tests of the next D = 2 / 3 / 4 levels
      // below the slot's node pair (2 + 4 + ... tests, every lane walking its own path of child
      // choices), and the group then replays the reference's decisions level by level along the
      // path actually taken.  A child test is a pure function of (node pair, transform, poses) and
      // the distance cannot change before the next leaf, so the replay is exact; tests off the
      // taken path are discarded and not counted.
      const int G = nT >= 5 ? 2 : (nT >= 3 ? 8 : (nT == 2 ? 16 : 32));
      const int D = nT >= 5 ? 1 : (nT >= 3 ? 2 : (nT == 2 ? 3 : 4));
      const int grp = lane / G, t = lane - grp * G;
      if (grp < nT)
      {
        const int slot = nth_slot(mT, (pass_no & 1) ? nT - 1 - grp : grp);  // alternate ends: no slot is starved
        const unsigned gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (grp * G));
        const int gbase = grp * G;
        double *stk = stack_base + (size_t)slot * stack_stride;
        int b1 = SI(I_CURB1, slot), b2 = SI(I_CURB2, slot);
        int sp = SI(I_SP, slot);
        const double dist = SD(F_DIST, slot), abs_err = SD(F_ABS, slot), rel_err = SD(F_REL, slot), upbound = SD(F_UPB, slot);
        double mint = SD(F_MINT, slot);
        double R[9], T[3];
        const bool from_cur = b1 >= 0;
        if (from_cur)
        {
#pragma unroll
          for (int i = 0; i < 9; i++) R[i] = SD(F_CUR + i, slot);
#pragma unroll
          for (int i = 0; i < 3; i++) T[i] = SD(F_CUR + 9 + i, slot);
        }
        else
        {
          // pop until an entry passes the descend test with the CURRENT distance (C2A.cpp:1281-1351);
          // entries that fail contribute their BV-level step bound.  All lanes of the group do this
          // redundantly on identical data.
          while (sp > 0)
          {
            const double *e = stk + (size_t)(sp - 1) * ENTRY_DOUBLES;
            sp--;
            const double2 dm = *reinterpret_cast<const double2 *>(e + 12);
            if (dm.y < upbound && ((dm.x < (dist - abs_err)) || (dm.x * (1 + rel_err) < dist)))
            {
              const double ids = e[14];
              b1 = __double2hiint(ids); b2 = __double2loint(ids);
#pragma unroll
              for (int i = 0; i < 6; i++)
              {
                const double2 v = *reinterpret_cast<const double2 *>(e + 2 * i);
                if (2 * i < 9) R[2 * i] = v.x; else T[2 * i - 9] = v.x;
                if (2 * i + 1 < 9) R[2 * i + 1] = v.y; else T[2 * i + 1 - 9] = v.y;
              }
              break;
            }
            if (dm.y < mint) mint = dm.y;
          }
        }

        if (b1 < 0)
        {
          // stack drained: this CA step's traversal is over
          if (t == 0) { SD(F_MINT, slot) = mint; SI(I_SP, slot) = 0; SI(I_STATE, slot) = ST_ADVANCE; }
        }
        else
        {
          NodeMeta ma, mb;
          if (from_cur)
          {
            ma.size = SD(F_CURSZ1, slot); ma.first_child = SI(I_CURFC1, slot);
            mb.size = SD(F_CURSZ2, slot); mb.first_child = SI(I_CURFC2, slot);
          }
          else { ma = load_meta(A.meta + b1); mb = load_meta(B.meta + b2); }
          if (ma.first_child < 0 && mb.first_child < 0)
          {
            if (t == 0)
            {
              SD(F_MINT, slot) = mint; SI(I_SP, slot) = sp; SI(I_CURB1, slot) = -1;
              SI(I_LEAFB1, slot) = b1; SI(I_LEAFB2, slot) = b2; SI(I_STATE, slot) = ST_LEAF;
            }
          }
          else
          {
            // ---- my test: level L (0-based), index i within the level; bit (L-l) of i picks the
            // child ('a' = 0, 'c' = 1) at level l of my walk
            const int ntests = (2 << D) - 2;
            const int L = 30 - __clz(t + 2);       // floor(log2(t+2)) - 1
            const int idx = t + 2 - (2 << L);
            bool valid = t < ntests;
            int n1 = b1, n2 = b2;
            NodeMeta cm1 = ma, cm2 = mb;           // NodeMeta of the node pair reached so far
            const double *gs = nullptr, *gt = nullptr, *rl = nullptr;
            if (valid)
            {
              for (int l = 0; l <= L; l++)
              {
                const bool l1 = cm1.first_child < 0, l2 = cm2.first_child < 0;
                if (l1 && l2) { valid = false; break; }  // my path runs through a leaf pair
                const int bit = (idx >> (L - l)) & 1;
                double Rc[9], Tc[3];
                if (l2 || (!l1 && (cm1.size > cm2.size)))
                {
                  // expansion of side 1, C2A.cpp:1194-1209
                  n1 = cm1.first_child + bit;
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
                  n2 = cm2.first_child + bit;
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
              }
            }
            // child BV test (C2A.cpp:1237-1276): RSS distance, direction to world frame, the two
            // directional motion bounds, the child's conservative step bound.  (R, T) now place my child pair.
            double d = 0.0, mt = 0.0;
            if (valid)
            {
              prefetch_l1(rl); prefetch_l1(rl + 8);  // R_loc is only consumed after the rectangle distance
              double S[3];
              // l(2) r ang_radius of both nodes: two 128-bit loads each
              const double2 la = __ldg(reinterpret_cast<const double2 *>(gs + 12)), ra2 = __ldg(reinterpret_cast<const double2 *>(gs + 14));
              const double2 lb = __ldg(reinterpret_cast<const double2 *>(gt + 12)), rb2 = __ldg(reinterpret_cast<const double2 *>(gt + 14));
              const double a0 = la.x, a1 = la.y, ra = ra2.x, e0 = lb.x, e1 = lb.y, rb = rb2.x;
              d = rss_rect_dist(R, T, a0, a1, e0, e1, S);
              d -= (ra + rb);
              d = (d < 0.0) ? 0.0 : d;
              if (d != 0.0)
              {
                double Rl[9], tmp[3], S1[3], S2[3], r1[9];
                load9v(Rl, rl);
                m_v(tmp, Rl, S);
#pragma unroll
                for (int i = 0; i < 9; i++) r1[i] = SD(F_R1 + i, slot);
                m_v(S1, r1, tmp);
                v_normalize(S1);  // S2 = S1 * -1 normalises to exactly -S1 (see c2a_motion.cuh)
                S2[0] = -S1[0]; S2[1] = -S1[1]; S2[2] = -S1[2];
                Motion m;  // only cv, axis, w are read by the bound
#pragma unroll
                for (int i = 0; i < 3; i++) { m.cv[i] = SD(F_CV1 + i, slot); m.axis[i] = SD(F_AX1 + i, slot); }
                m.w = SD(F_W1, slot);
                const double mb1 = motion_bound_bv_unit(m, ra2.y, S1);
#pragma unroll
                for (int i = 0; i < 3; i++) { m.cv[i] = SD(F_CV2 + i, slot); m.axis[i] = SD(F_AX2 + i, slot); }
                m.w = SD(F_W2, slot);
                const double mb2 = motion_bound_bv_unit(m, rb2.y, S2);
                mt = (d) / (mb1 + mb2);
                if (mt <= 0) mt = 0.0;
              }
            }
            const bool my_leafpair = valid && cm1.first_child < 0 && cm2.first_child < 0;
            const bool v = mt < upbound && ((d < (dist - abs_err)) || (d * (1 + rel_err) < dist));
            const int flags = (valid ? 1 : 0) | (my_leafpair ? 2 : 0) | (v ? 4 : 0);

            // ---- replay the reference's depth-first walk (C2A.cpp:1281-1351) over the evaluated tree: every lane
            // of the group computes the same walk from the shuffled (d, mint, flags).  When a node pair's two
            // children are both pruned the walk returns to the most recent pending far child; if that one was
            // pushed during THIS pass its own child tests are in the tree as well (the distance has not changed,
            // so it still passes the descend test it passed when pushed) and the walk carries on through it.
            // The pass ends at a leaf pair, at a pair whose children lie below the tree, or when nothing
            // pushed in this pass is pending.
            int P = 0, l = 0;      // expanding the pair reached by test (l - 1, P); the slot's node pair for l = 0
            int levels = 0;        // expansions committed
            int write_cur = -1;    // group lane whose child becomes the current entry
            int write_leaf = -1;   // group lane whose child is a leaf pair to hand to the LEAF phase
            bool me_push = false; int my_push_sp = 0;
            unsigned pend = 0; int npend = 0;  // far children pushed in this pass and still pending (lane ids, 8 bits each)
            if (D == 1)
            {
              // the steady state (a lane pair per slot): one expansion, no walk
              const double d_o = __shfl_xor_sync(gmask, d, 1), m_o = __shfl_xor_sync(gmask, mt, 1);
              const int f_o = __shfl_xor_sync(gmask, flags, 1);
              const double d_a = t ? d_o : d, d_c = t ? d : d_o, m_a = t ? m_o : mt, m_c = t ? mt : m_o;
              const int f_a = t ? f_o : flags, f_c = t ? flags : f_o;
              levels = 1;
              const bool v_a = f_a & 4, v_c = f_c & 4;
              const bool c_first = d_c < d_a;  // ties visit 'a' first (the test is d2 < d1)
              const bool v_near = c_first ? v_c : v_a, v_far = c_first ? v_a : v_c;
              const int near_lane = c_first ? 1 : 0, far_lane = c_first ? 0 : 1;
              if (!v_a && m_a < mint) mint = m_a;
              if (!v_c && m_c < mint) mint = m_c;
              if (v_near || v_far)
              {
                const int next_lane = v_near ? near_lane : far_lane;
                if (v_near && v_far)
                {
                  if (t == far_lane) { me_push = true; my_push_sp = sp; }
                  sp++;
                }
                if ((next_lane ? f_c : f_a) & 2) write_leaf = next_lane;
                else write_cur = next_lane;
              }
            }
            else
            while (true)
            {
              const int la = (2 << l) - 2 + 2 * P, lc = la + 1;  // group lanes of children 'a' and 'c'
              const double d_a = __shfl_sync(gmask, d, gbase + la), d_c = __shfl_sync(gmask, d, gbase + lc);
              const double m_a = __shfl_sync(gmask, mt, gbase + la), m_c = __shfl_sync(gmask, mt, gbase + lc);
              const int f_a = __shfl_sync(gmask, flags, gbase + la), f_c = __shfl_sync(gmask, flags, gbase + lc);
              // (both children of an inner node exist together, so f_a and f_c are valid together)
              levels++;
              const bool v_a = f_a & 4, v_c = f_c & 4;
              const bool c_first = d_c < d_a;  // ties visit 'a' first (the test is d2 < d1)
              const bool v_near = c_first ? v_c : v_a, v_far = c_first ? v_a : v_c;
              const int near_lane = c_first ? lc : la, far_lane = c_first ? la : lc;
              if (!v_a && m_a < mint) mint = m_a;
              if (!v_c && m_c < mint) mint = m_c;
              int next_lane, f_next;
              if (v_near)
              {
                next_lane = near_lane;
                if (v_far)
                {
                  if (t == far_lane) { me_push = true; my_push_sp = sp; }
                  sp++;
                  pend |= (unsigned)far_lane << (8 * npend);
                  npend++;
                }
                f_next = (next_lane == la) ? f_a : f_c;
              }
              else if (v_far) { next_lane = far_lane; f_next = (next_lane == la) ? f_a : f_c; }
              else
              {
                if (npend == 0) { write_cur = -1; break; }
                npend--;
                next_lane = (int)((pend >> (8 * npend)) & 0xffu);
                pend &= ~(0xffu << (8 * npend));
                sp--;
                if (t == next_lane) me_push = false;
                f_next = __shfl_sync(gmask, flags, gbase + next_lane);
              }
              if (f_next & 2) { write_leaf = next_lane; write_cur = -1; break; }
              const int Ln = 30 - __clz(next_lane + 2);  // level of that test
              if (Ln + 1 >= D) { write_cur = next_lane; break; }
              l = Ln + 1;
              P = next_lane + 2 - (2 << Ln);
            }

            if (me_push)
            {
              double *e = stk + (size_t)my_push_sp * ENTRY_DOUBLES;
              double2 *e2 = reinterpret_cast<double2 *>(e);
              e2[0] = make_double2(R[0], R[1]); e2[1] = make_double2(R[2], R[3]);
              e2[2] = make_double2(R[4], R[5]); e2[3] = make_double2(R[6], R[7]);
              e2[4] = make_double2(R[8], T[0]); e2[5] = make_double2(T[1], T[2]);
              e2[6] = make_double2(d, mt); e2[7] = make_double2(__hiloint2double(n1, n2), 0.0);
            }
            if (t == write_cur)
            {
#pragma unroll
              for (int i = 0; i < 9; i++) SD(F_CUR + i, slot) = R[i];
#pragma unroll
              for (int i = 0; i < 3; i++) SD(F_CUR + 9 + i, slot) = T[i];
              SI(I_CURB1, slot) = n1; SI(I_CURB2, slot) = n2;
              SD(F_CURSZ1, slot) = cm1.size; SI(I_CURFC1, slot) = cm1.first_child;
              SD(F_CURSZ2, slot) = cm2.size; SI(I_CURFC2, slot) = cm2.first_child;
              // warm L1 with the node pair the next expansion of this slot will fetch
              const bool nl1 = cm1.first_child < 0, nl2 = cm2.first_child < 0;
              const double *nx = (nl2 || (!nl1 && (cm1.size > cm2.size))) ? A.geom + (size_t)cm1.first_child * GEOM_STRIDE
                                                                          : B.geom + (size_t)cm2.first_child * GEOM_STRIDE;
              prefetch_l1(nx); prefetch_l1(nx + GEOM_STRIDE);
            }
            if (t == write_leaf)
            {
              SI(I_LEAFB1, slot) = n1; SI(I_LEAFB2, slot) = n2; SI(I_STATE, slot) = ST_LEAF;
            }
            if (t == 0)
            {
              if (SOLVE_STATS_PTR && G == 32) { atomicAdd(SOLVE_STATS_PTR + 9, 1ull); atomicAdd(SOLVE_STATS_PTR + 10, (unsigned long long)levels); }
              SD(F_MINT, slot) = mint;
              SI(I_SP, slot) = sp;
              SI(I_NBV, slot) = SI(I_NBV, slot) + 2 * levels;
              if (write_cur < 0) SI(I_CURB1, slot) = -1;
            }
          }
        }
      }
    }
    else if (phase == ST_LEAF)
    {
      // ---------------------------------------------------------------------- LEAF -----------
      // C2A.cpp:1141-1183
      if (lane < nL)
      {
        const int slot = nth_slot(mL, lane);
        const int b1 = SI(I_LEAFB1, slot), b2 = SI(I_LEAFB2, slot);
        double Rrel[9], Trel[3], p[3], qq[3];
#pragma unroll
        for (int i = 0; i < 9; i++) Rrel[i] = SD(F_RREL + i, slot);
#pragma unroll
        for (int i = 0; i < 3; i++) Trel[i] = SD(F_TREL + i, slot);
        const int ta = -A.meta[b1].first_child - 1, tb = -B.meta[b2].first_child - 1;
        const double dTri = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * ta, B.tris + (size_t)TRI_STRIDE * tb, p, qq);
        if (dTri <= SD(F_DIST, slot))
        {
          SD(F_DIST, slot) = dTri;
          double r1[9], tt1[3], w1[3], w2[3], S1[3], S2[3], tmp[3];
#pragma unroll
          for (int i = 0; i < 9; i++) r1[i] = SD(F_R1 + i, slot);
#pragma unroll
          for (int i = 0; i < 3; i++) tt1[i] = SD(F_TT1 + i, slot);
          m_v(tmp, r1, p); v_add(w1, tmp, tt1);
          m_v(tmp, r1, qq); v_add(w2, tmp, tt1);
          v_sub(S1, w2, w1);
          v_normalize(S1);  // S2 = S1 * -1 normalises to exactly -S1 (see c2a_motion.cuh)
          S2[0] = -S1[0]; S2[1] = -S1[1]; S2[2] = -S1[2];
          if (args.out.p1p2)
          {
            const long long q = ((long long)SI(I_QHI, slot) << 32) | (unsigned)SI(I_QLO, slot);
#pragma unroll
            for (int i = 0; i < 3; i++) { args.out.p1p2[6 * q + i] = p[i]; args.out.p1p2[6 * q + 3 + i] = qq[i]; }
          }
          Motion m;
#pragma unroll
          for (int i = 0; i < 3; i++) { m.cv[i] = SD(F_CV1 + i, slot); m.axis[i] = SD(F_AX1 + i, slot); }
          m.w = SD(F_W1, slot);
          const double mb1 = motion_bound_leaf_unit(m, __ldg(A.geom + (size_t)b1 * GEOM_STRIDE + 15), S1);
#pragma unroll
          for (int i = 0; i < 3; i++) { m.cv[i] = SD(F_CV2 + i, slot); m.axis[i] = SD(F_AX2 + i, slot); }
          m.w = SD(F_W2, slot);
          const double mb2 = motion_bound_leaf_unit(m, __ldg(B.geom + (size_t)b2 * GEOM_STRIDE + 15), S2);
          double mt = (dTri) / (mb1 + mb2);
          if (mt < 0.0) mt = 0.0;
          if (mt <= SD(F_MINT, slot)) SD(F_MINT, slot) = mt;
          SI(I_LASTA, slot) = ta; SI(I_LASTB, slot) = tb;  // o1->last_tri = t1; o2->last_tri = t2 (C2A.cpp:1175-1176)
        }
        SI(I_NTRI, slot) = SI(I_NTRI, slot) + 1;
        SI(I_STATE, slot) = ST_TRAVERSE;
      }
    }
    else
    {
      // -------------------------------------------------------------------- ADVANCE ----------
      // CA-loop bookkeeping after a finished step / result write-out / claim / next step's set-up
      if (lane < nA)
      {
        const int slot = nth_slot(mA, lane);
        long long q = ((long long)SI(I_QHI, slot) << 32) | (unsigned)SI(I_QLO, slot);
        bool pending = SI(I_PENDING, slot) != 0;
        int numCA = SI(I_NUMCA, slot);
        double lamda = SD(F_LAMDA, slot);
        bool handed_over = false;
        if (q >= 0 && !pending && args.step_in)
        {
          // single-step mode: report this traversal and release the slot
          const c2a_b200_results &o = args.out;
          if (o.status) o.status[q] = C2A_B200_QUERY_OK;
          if (o.num_bv_tests) o.num_bv_tests[q] = SI(I_NBV, slot);
          if (o.num_tri_tests) o.num_tri_tests[q] = SI(I_NTRI, slot);
          if (o.distance) o.distance[q] = SD(F_DIST, slot);
          if (o.mint) o.mint[q] = SD(F_MINT, slot);
          if (o.last_tri) { o.last_tri[2 * q] = SI(I_LASTA, slot); o.last_tri[2 * q + 1] = SI(I_LASTB, slot); }
          q = -1;
        }
        bool just_continued = false;
        if (q >= 0 && !pending)
        {
          // a step just ended: C2A_QueryTimeOfContact's loop, C2A.cpp:2053-2123
          const double *rec = args.motions + (size_t)(2 * MOTION_DOUBLES) * q;
          const double dist = SD(F_DIST, slot), mint = SD(F_MINT, slot);
          double lastLamda = SD(F_LASTL, slot);
          int nItrs = SI(I_NITRS, slot);
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
                else
                {
                  lastLamda = lamda;
                  numCA++;
                  SD(F_UPB, slot) = 1.0 - lamda;
                  pending = true;
                  just_continued = true;
                }
              }
            }
          }
          SI(I_NITRS, slot) = nItrs;
          SD(F_LASTL, slot) = lastLamda;
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
            if (o.num_bv_tests) o.num_bv_tests[q] = SI(I_NBV, slot);
            if (o.num_tri_tests) o.num_tri_tests[q] = SI(I_NTRI, slot);
            if (o.toc) o.toc[q] = toc;
            if (o.distance) o.distance[q] = dist;
            if (o.mint) o.mint[q] = mint;
            if (o.last_tri) { o.last_tri[2 * q] = SI(I_LASTA, slot); o.last_tri[2 * q + 1] = SI(I_LASTB, slot); }
            if (SOLVE_TRACE_PTR) SOLVE_TRACE_PTR[2 * q + 1] = global_ns();
            q = -1;
          }
        }

        // hand-over to the wide kernel: the queue is empty, the warp is running out of queries, and this query's
        // remaining steps are all in exact mode (numCA > 5, C2A.cpp:1869) -> write its CA-loop state out and retire
        if (args.spill_recs && just_continued && numCA > 5 && n_live <= args.spill_live &&
            *(volatile unsigned long long *)args.counter >= (unsigned long long)args.n)
        {
          // three lists, claimed in this order by the wide kernel: the more steps a query has taken, the more it is
          // likely to have left (longest-expected first)
          const int bucket = numCA >= SPILL_CA_HEAVY ? 0 : (numCA >= SPILL_CA_MID ? 1 : 2);
          const unsigned long long idx = atomicAdd(args.spill_count + bucket, 1ull);
          if (idx < (unsigned long long)args.spill_cap)
          {
            double *r = args.spill_recs + ((size_t)bucket * args.spill_cap + idx) * MB_DOUBLES;
            r[1] = lamda; r[2] = SD(F_LASTL, slot); r[3] = SD(F_MINT, slot); r[4] = SD(F_UPB, slot);
            r[5] = __hiloint2double(numCA, SI(I_NITRS, slot)); r[6] = __hiloint2double(SI(I_NBV, slot), SI(I_NTRI, slot));
            r[7] = __hiloint2double(SI(I_LASTA, slot), SI(I_LASTB, slot));
            // the wide kernel may be running beside this one: the query index (pre-set to -1 by the host) goes last and
            // marks the record as complete
            __threadfence();
            *reinterpret_cast<volatile long long *>(r) = q;
            SI(I_STATE, slot) = ST_EXIT;
            q = -1; pending = false; handed_over = true;
          }
        }

        if (q < 0 && !handed_over)
        {
          // claim the next query
          const long long nq = (long long)atomicAdd(args.counter, 1ull);
          if (nq >= args.n)
          {
            SI(I_STATE, slot) = ST_EXIT;
            if (SOLVE_STATS_PTR) { atomicMin(SOLVE_STATS_PTR + 7, global_ns()); atomicMax(SOLVE_STATS_PTR + 8, global_ns()); }  // batch drained / last slot retired
          }
          else
          {
            q = args.order ? (long long)__ldg(args.order + nq) : nq;
            if (SOLVE_TRACE_PTR) SOLVE_TRACE_PTR[2 * q] = global_ns();
            const double *rec = args.motions + (size_t)(2 * MOTION_DOUBLES) * q;
            const double w1 = __ldg(rec + 18), w2 = __ldg(rec + MOTION_DOUBLES + 18);
            if (!args.step_in && w1 < 1e-8 && w2 < 1e-8)
            {
              // translation-only branch of the reference (C2A.cpp:2391-2395): solved by c2a_translation_kernel,
              // which the host launches over the same batch right after this kernel
              q = -1;  // stay in ADVANCE: claim another one next round
            }
            else
            {
#pragma unroll
              for (int i = 0; i < 3; i++)
              {
                SD(F_CV1 + i, slot) = __ldg(rec + 12 + i); SD(F_AX1 + i, slot) = __ldg(rec + 15 + i);
                SD(F_CV2 + i, slot) = __ldg(rec + MOTION_DOUBLES + 12 + i); SD(F_AX2 + i, slot) = __ldg(rec + MOTION_DOUBLES + 15 + i);
              }
              SD(F_W1, slot) = w1; SD(F_W2, slot) = w2;
              SI(I_SEEDA, slot) = seed_or_zero(args.seedA, q, A.n_tris);
              SI(I_SEEDB, slot) = seed_or_zero(args.seedB, q, B.n_tris);
              numCA = 0; lamda = 0;
              SI(I_NITRS, slot) = 0; SI(I_NBV, slot) = 0; SI(I_NTRI, slot) = 0;
              SI(I_LASTA, slot) = -1; SI(I_LASTB, slot) = -1;
              SD(F_LASTL, slot) = 0; SD(F_UPB, slot) = 1; SD(F_MINT, slot) = 1; SD(F_DIST, slot) = 0;
              if (args.step_in)
              {
                const double *si_ = args.step_in + (size_t)STEP_IN_DOUBLES * q;
                numCA = (int)__ldg(si_ + 24); SD(F_MINT, slot) = __ldg(si_ + 25); SD(F_UPB, slot) = __ldg(si_ + 26);
              }
              if (args.out.p1p2)
              {
                // res->p1/p2 stay zero unless a leaf improves the distance (the LEAF phase writes them in place)
#pragma unroll
                for (int i = 0; i < 6; i++) args.out.p1p2[6 * q + i] = 0.0;
              }
              pending = true;
            }
          }
          SI(I_QLO, slot) = (int)(unsigned)(q & 0xffffffffll); SI(I_QHI, slot) = (int)(q >> 32);
        }

        if (q >= 0 && pending)
        {
          // C2A_TimeOfContactStep, C2A.cpp:1791-1894
          const double *rec = args.motions + (size_t)(2 * MOTION_DOUBLES) * q;
          double r1[9], tt1[3], R2[9], T2[3], Tt[3], Rt[9], g1[12], g2[12], R[9], T[3], Rrel[9], Trel[3];
          if (args.step_in)
          {
            const double *si_ = args.step_in + (size_t)STEP_IN_DOUBLES * q;
            load9(r1, si_); load3(tt1, si_ + 9); load9(R2, si_ + 12); load3(T2, si_ + 21);
          }
          else if (numCA == 0)
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
          const double dist = tri_distance_v(Rrel, Trel, A.tris + (size_t)TRI_STRIDE * SI(I_SEEDA, slot),
                                             B.tris + (size_t)TRI_STRIDE * SI(I_SEEDB, slot), p, qq);
          double mint = SD(F_MINT, slot);
          if (numCA == 0) mint = 1;
          if (mint <= 0.005 || dist <= 0.5 || numCA > 5) { SD(F_ABS, slot) = 0; SD(F_REL, slot) = 0; }
          else { SD(F_ABS, slot) = 1e+30; SD(F_REL, slot) = (numCA <= 2) ? 3 : 0.5; }
          SD(F_MINT, slot) = 1;
          SD(F_DIST, slot) = dist;
#pragma unroll
          for (int i = 0; i < 9; i++) { SD(F_R1 + i, slot) = r1[i]; SD(F_RREL + i, slot) = Rrel[i]; SD(F_CUR + i, slot) = R[i]; }
#pragma unroll
          for (int i = 0; i < 3; i++) { SD(F_TT1 + i, slot) = tt1[i]; SD(F_TREL + i, slot) = Trel[i]; SD(F_CUR + 9 + i, slot) = T[i]; }
          // the root pair is descended unconditionally
          SI(I_CURB1, slot) = 0; SI(I_CURB2, slot) = 0; SI(I_SP, slot) = 0;
          {
            const NodeMeta ra = A.meta[0], rb = B.meta[0];
            SD(F_CURSZ1, slot) = ra.size; SI(I_CURFC1, slot) = ra.first_child;
            SD(F_CURSZ2, slot) = rb.size; SI(I_CURFC2, slot) = rb.first_child;
          }
          pending = false;
          SI(I_STATE, slot) = ST_TRAVERSE;
        }
        if (handed_over) { SI(I_QLO, slot) = -1; SI(I_QHI, slot) = -1; }
        SI(I_NUMCA, slot) = numCA;
        SD(F_LAMDA, slot) = lamda;
        SI(I_PENDING, slot) = pending ? 1 : 0;
      }
    }
    __syncwarp();
    if (SOLVE_STATS_PTR && lane == 0 && steady) atomicAdd(SOLVE_STATS_PTR + 11 + (phase == ST_TRAVERSE ? 0 : (phase == ST_LEAF ? 1 : 2)), (unsigned long long)(clock64() - pass_t0));
    if (SOLVE_STATS_PTR && lane == 0 && n_live == 1)
    {
      // a query alone on its warp (the tail of a launch, or a single-query call): passes and cycles per phase
      const int k = 14 + 2 * (phase == ST_TRAVERSE ? 0 : (phase == ST_LEAF ? 1 : 2));
      atomicAdd(SOLVE_STATS_PTR + k, 1ull); atomicAdd(SOLVE_STATS_PTR + k + 1, (unsigned long long)(clock64() - pass_t0));
    }
  }
  // this warp hands nothing over any more (its records are complete: fence before the count)
  if (args.spill_count && lane == 0) { __threadfence(); atomicAdd(args.spill_count + SPILL_CTL_DONE, 1ull); }
#undef SD
#undef SI
}

}  // namespace c2a
