/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * CPU oracle ("port"): a plain, scalar FP64 restatement of the reference's
 * controlled conservative-advancement CCD hot path
 *   C2A_Solve -> C2A_QueryTimeOfContact -> C2A_TimeOfContactStep -> TOCStepRecurse_Dis
 * (reference: /root/reference/C2A/src/C2A.cpp, InterpMotion.cpp, C2A_RectDist.h,
 * LinearMath.h; each function in c2a_oracle.cpp cites the file:line it follows).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may call this.  The product (c2a_b200/csrc) never links it.
 *
 * Pinning: the reference ships no golden vectors, and its
 * PQP dependency is absent, so this port is pinned against the reference's own
 * object code run here: oracle/_ref (the reference's five .cpp compiled verbatim
 * against oracle/pqp_shim) -- bit-exact on every query of the fixtures under
 * tests/golden/ (see oracle/README.md).  Parity at the PQP boundary itself
 * (Meigen, TriDist; for the discrete queries also obb_disjoint, TriContact and
 * the BVTQ queue's tie order) is "unpinned" in the sense of SURVEY.md section
 * 8c: PQP is unpinned upstream and absent here, those functions are restated
 * once in oracle/pqp_shim; TriDist is cross-checked bit-exactly against the
 * reference's in-tree copy (C2A/src/C2A.cpp:165-405).
 *
 * Build: g++ -O2 -ffp-contract=off (no -ffast-math).
 */
#ifndef C2A_ORACLE_H
#define C2A_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened RSS-BVH: the hot fields of C2A_BV (C2A/C2A_BV.h:33-77 + PQP BV) and
 * the triangles (PQP Tri p1,p2,p3), as produced by the reference's builder. */
typedef struct orc_bvh
{
  int32_t n_nodes, n_tris;
  const double *R;            /* [n_nodes][9]  BV::R, parent-relative (root: model frame) */
  const double *Tr;           /* [n_nodes][3]  BV::Tr, parent-relative                    */
  const double *l;            /* [n_nodes][2]  rectangle side lengths                     */
  const double *r;            /* [n_nodes]     RSS radius                                 */
  const double *R_loc;        /* [n_nodes][9]  C2A_BV::R_loc, model frame                 */
  const double *ang_radius;   /* [n_nodes]     C2A_BV::angularRadius                      */
  const int32_t *first_child; /* [n_nodes]     >=0 child index; <0: -(tri+1)              */
  const double *tris;         /* [n_tris][9]   p1,p2,p3 in builder (permuted) order       */
} orc_bvh;

typedef struct orc_result
{
  int32_t collisionfree;  /* C2A_TimeOfContactResult::collisionfree */
  int32_t numCA;          /* ::numCA  (== number_of_iteration of C2A_Solve) */
  int32_t num_bv_tests;   /* ::num_bv_tests */
  int32_t num_tri_tests;  /* ::num_tri_tests */
  double toc;             /* ::toc */
  double distance;        /* ::distance */
  double mint;            /* ::mint of the last step */
  double p1[3], p2[3];    /* ::p1, ::p2 (model-1 frame) */
  double pose_toc[24];    /* C2A_Solve's trans0, trans1 as R(9)+T(3) each; written only on a hit */
  int32_t last_tri_a, last_tri_b; /* o1->last_tri / o2->last_tri as the traversal leaves them (C2A.cpp:1175-1176):
                                     triangle indices of the last leaf that improved the distance, -1 if none */
} orc_result;

/* geometry kernels */
double orc_rect_dist(const double Rab[9], const double Tab[3], const double a[2], const double b[2],
                     double P[3], double Q[3], double S[3]);
double orc_tri_dist(double P[3], double Q[3], const double S[9], const double T[9]);
void orc_seg_points(double VEC[3], double X[3], double Y[3], const double P[3], const double A[3],
                    const double Q[3], const double B[3]);
double orc_tri_distance(const double R[9], const double T[3], const double t1[9], const double t2[9],
                        double p[3], double q[3]);

/* motion (CInterpMotion_Linear) */
typedef struct orc_motion
{
  double Rs[9], Ts[3], Re[9], Te[3]; /* transform_s, transform_t */
  double cv[3], axis[3], ang_vel;    /* cv, m_axis, m_angVel */
  double Rc[9], Tc[3];               /* transform (current pose; mutated by integrate) */
} orc_motion;
void orc_motion_init(orc_motion *m, const double R0[9], const double T0[3], const double R1[9],
                     const double T1[3]);
void orc_motion_integrate(orc_motion *m, double t, double q_out[4] /* x,y,z,w; may be NULL */);
double orc_motion_bound_bv(const orc_motion *m, double ang_radius, double N[3]);
double orc_motion_bound_leaf(const orc_motion *m, double ang_radius, double S[3]);

/* One query.  poses = trans00, trans01, trans10, trans11 as R(9 row-major)+T(3) each (48 doubles).
 * seedA/seedB: triangle indices playing res->last_triA/B.
 * Mirrors C2A_Solve minus the contact pass when tol_d = tol_t = 1e-4 (C2A.cpp:2384-2385). */
void orc_solve(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA,
               int32_t seedB, double tol_d, double tol_t, orc_result *out);

/* ---- contact pass of C2A_Solve (C2A_QueryContact, C2A/src/C2A.cpp:1937-1966) ---------------------- */
typedef struct orc_contact
{
  int32_t type_a, type_b;   /* ContactF::FeatureType_A/B: 1 vertex, 2 edge, 3 face */
  int32_t fid_a[3], fid_b[3]; /* vertex indices of the feature (entries the reference leaves uninitialised are -1) */
  int32_t tri_a, tri_b;     /* triangle indices (builder order) */
  double pa[3], pb[3];      /* closest points: on A in A's frame, on B in B's frame */
  double dist;
} orc_contact;

/* the reference's in-tree TriDist with contact features, C2A/src/C2A.cpp:165-405 */
double orc_tri_dist_features(double P[3], double Q[3], const double S[9], const double T[9], int32_t *f1_type,
                             int32_t *f1_fid, int32_t *f2_type, int32_t *f2_fid, int32_t *collided);

/* All triangle pairs within `threshold` at the given poses (R(9)+T(3) each), in the reference's visiting
 * order (the reference push_front()s them, so its list is this order reversed).  vidx_a/b: [n_tris][3]
 * vertex indices per triangle (builder order) or NULL.  Returns the number found; at most max_out are
 * written. */
int64_t orc_contacts(const orc_bvh *A, const orc_bvh *B, const int32_t *vidx_a, const int32_t *vidx_b,
                     const double pose1[12], const double pose2[12], double threshold, int64_t max_out,
                     orc_contact *out);

/* Batch over n queries on n_threads std::threads (static interleave). */
void orc_solve_batch(const orc_bvh *A, const orc_bvh *B, const double *poses, int64_t n,
                     const int32_t *seedA, const int32_t *seedB, double tol_d, double tol_t,
                     orc_result *out, int32_t n_threads);

/* C2A_Distance (C2A/src/C2A_PQP.cpp:970-1056), depth-first routine (qsize <= 2).  tri_a / tri_b: builder-order
 * indices of the closest triangle pair (the reference reports Tri::id and leaves them in o->last_tri). */
typedef struct orc_distance_result
{
  double distance;
  double p1[3], p2[3]; /* closest points, each in its own model's frame */
  int32_t tri_a, tri_b;
  int32_t num_bv_tests, num_tri_tests;
} orc_distance_result;
void orc_distance(const orc_bvh *A, const orc_bvh *B, const double pose24[24], int32_t seedA, int32_t seedB,
                  double rel_err, double abs_err, orc_distance_result *res);

/* C2A_Distance with the priority-queue routine it takes for qsize > 2 (C2ADistanceQueueRecurse, C2A/src/C2A_PQP.cpp:624-787);
 * among equally distant pending pairs the one queued first is taken first, as oracle/pqp_shim/BVTQ.h does. */
void orc_distance_queue(const orc_bvh *A, const orc_bvh *B, const double pose24[24], int32_t seedA, int32_t seedB,
                        double rel_err, double abs_err, int32_t qsize, orc_distance_result *res);

/* C2A_Collide, both overloads (C2A/src/C2A_PQP.cpp:910-968, 1199-1280).  dA/ToA/dB/ToB: [n_nodes][3] OBB half-dimensions
 * and centres (BV::d, BV::To).  orc_obb_disjoint / orc_tri_contact restate PQP's un-vendored obb_disjoint / TriContact in
 * the arithmetic of oracle/pqp_shim (what the compiled reference links). */
int32_t orc_obb_disjoint(const double B[9], const double T[3], const double a[3], const double b[3]);
int32_t orc_tri_contact(const double P[9], const double Q[9]);
int32_t orc_collide(const orc_bvh *A, const orc_bvh *B, const double *dA, const double *ToA, const double *dB,
                    const double *ToB, const double pose24[24], int32_t flag, int32_t max_pairs, int32_t *pairs,
                    int32_t *num_bv_tests, int32_t *num_tri_tests);
void orc_collide_distance(const orc_bvh *A, const orc_bvh *B, const double *dA, const double *dB, const double pose24[24],
                          int32_t seedA, int32_t seedB, double rel_err, double abs_err, orc_distance_result *res);

/* Round-2 design study (test infrastructure): a CA step split into subtrees run under a guessed entry distance with a
 * recorded validity interval, stitched back in the reference's order.  Work is counted in node-pair visits. */
typedef struct orc_spec_stats
{
  long long steps, tasks, reached, valid;      /* exact-mode steps, frontier subtrees, those the sequential order reaches, those whose guess held */
  long long work_seq, work_fallback, work_wasted, work_max_task, work_top;
  double par_time[3];                          /* estimated critical path with 8 / 32 / 128 workers */
  int32_t depth;
} orc_spec_stats;
void orc_solve_spec(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB, double tol_d,
                    double tol_t, int32_t K, orc_result *out, orc_spec_stats *stats);

/* Round-2 design study: every CA step replayed over the previous step's visit list (records re-evaluated for the
 * current poses before the walk, misses evaluated on the spot).  Counts are node-pair visits. */
typedef struct orc_replay_stats
{
  long long steps, visits, hits, seq_hits, misses; /* walk: visits = hits + misses; seq_hits = hits on the next record of the list */
  long long preeval, wasted;                       /* records evaluated before the walks; of those, never visited */
} orc_replay_stats;
void orc_solve_replay(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB, double tol_d,
                      double tol_t, orc_result *out, orc_replay_stats *stats);

/* Round-2 algorithm (the CPU statement of c2a_b200/csrc/c2a_wide.cuh): exact-mode CA steps as a depth-first traversal
 * that pops `window` node pairs per round, resolves the distance updates in key order and folds at the end. */
typedef struct orc_wide_stats
{
  long long window;                                 /* in: node pairs popped per round */
  long long leaf_batch;                             /* in: 0 = leaf pairs are tested in the window they are popped in; n > 0 = they wait for a
                                                       LEAF pass, run when n are waiting or nothing else is left (32 per pass), as on the device */
  long long leaf_passes;
  long long steps, redo, anomalies, closure_fail, events, rounds, max_width, max_stack, max_unresolved;
  long long wide_tests, wide_tests_visited, wide_leaves, wide_leaves_visited;
} orc_wide_stats;
void orc_solve_wide(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB, double tol_d,
                    double tol_t, orc_result *out, orc_wide_stats *stats);

int64_t orc_solve_visits(const orc_bvh *A, const orc_bvh *B, const double poses[48], int32_t seedA, int32_t seedB, double tol_d,
                         double tol_t, orc_result *out, uint64_t *visits, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif
