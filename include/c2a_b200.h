/* c2a_b200 -- C ABI of the B200-native C2A continuous-collision-detection hot path.
 *
 * Drop-in boundary.  The reference (EwhaGlab/C2A) has no FFI layer: its boundary is the C++ header
 * C2A/C2A.h.  The C++ shims in include/C2A/ keep those entry points (C2A_Solve C2A/C2A.h:23-35,
 * C2A_QueryTimeOfContact C2A/C2A.h:274-281, C2A_TimeOfContactStep C2A/src/C2A.cpp:1778-1789) and
 * call the functions below; a batched C2A_SolveBatch is added.  Everything here is extern "C",
 * plain pointers and sizes; no C++ or torch types cross it.  See INTEGRATION.md.
 *
 * All functions return 0 on success, a negative C2A_B200_ERR_* code otherwise;
 * c2a_b200_last_error() describes the last failure on the calling thread.  Nothing here prints
 * to stdout or calls exit() (the reference does both: C2A/src/C2A.cpp:2408, C2A/LinearMath.h:768).
 * There is no CPU fallback: without a CUDA device every compute entry fails with
 * C2A_B200_ERR_CUDA.
 */
#ifndef C2A_B200_H
#define C2A_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2A_B200_OK 0
#define C2A_B200_ERR_ARG (-1)         /* null pointer, bad size, malformed BVH */
#define C2A_B200_ERR_CUDA (-2)        /* CUDA runtime error (including: no device) */
#define C2A_B200_ERR_DEPTH (-3)       /* depth(A)+depth(B)+2 > 512: beyond what the batched CCD kernel allocates per query slot */
#define C2A_B200_ERR_DEVICE (-4)      /* models live on different devices */

/* per-query status written to c2a_b200_results.status */
#define C2A_B200_QUERY_OK 0
#define C2A_B200_QUERY_TRANSLATION_ONLY 1 /* (round 1 only: translation-only queries on hierarchies deeper than that branch's
                                             local stack.  No longer reported: deep hierarchies run with their stacks in
                                             global memory, so every query of a batch ends with status 0) */

/* Flattened RSS bounding-volume hierarchy of one C2A_Model after EndModel(): the hot fields of
 * C2A_BV (C2A/C2A_BV.h:33-77 on top of PQP's BV) and the triangles (PQP Tri p1,p2,p3) in the
 * builder's (permuted) order.  Children of node n are first_child[n] and first_child[n]+1
 * (C2A/src/C2A_Build.cpp:456-457); first_child[n] < 0 marks a leaf holding triangle
 * -first_child[n]-1 (:451).  R/Tr are parent-relative (:533-538), R_loc is in the model frame
 * (:439).  Host pointers; the upload copies everything. */
typedef struct c2a_b200_bvh
{
  int32_t n_nodes;
  int32_t n_tris;
  const double *R;            /* [n_nodes][9] row-major */
  const double *Tr;           /* [n_nodes][3] */
  const double *l;            /* [n_nodes][2] */
  const double *r;            /* [n_nodes]    */
  const double *R_loc;        /* [n_nodes][9] row-major */
  const double *ang_radius;   /* [n_nodes]    C2A_BV::angularRadius */
  const int32_t *first_child; /* [n_nodes]    */
  const double *tris;         /* [n_tris][9]  p1,p2,p3 */
  const int32_t *tri_vidx;    /* [n_tris][3]  vertex indices of each triangle (C2A_Tri::Index(), C2A/C2A_Tri.h:46-52),
                                              builder order; only used to label contact features; may be NULL */
  const double *obb_d;        /* [n_nodes][3] OBB half-dimensions BV::d (C2A/src/C2A_BV.cpp:414-416) and           */
  const double *obb_To;       /* [n_nodes][3] OBB centre BV::To, parent-relative (C2A_Build.cpp:540-541): read by   */
                              /*              the C2A_Collide entries only; both may be NULL (those entries then    */
                              /*              refuse the model)                                                     */
} c2a_b200_bvh;

/* One contact of the contact pass (ContactF, C2A/C2A.h:117-153). */
typedef struct c2a_b200_contact
{
  int32_t type_a, type_b;     /* FeatureType_A/B: 1 vertex, 2 edge, 3 face */
  int32_t fid_a[3], fid_b[3]; /* FeatureID_A/B: vertex indices of the feature; entries the reference leaves
                                 uninitialised (2nd/3rd of a vertex, 3rd of an edge) are -1; all -1 without tri_vidx */
  int32_t tri_a, tri_b;       /* TriangleID_A/B (builder order) */
  double pa[3], pb[3];        /* P_A in model A's frame, P_B in model B's frame */
  double dist;                /* Distance */
} c2a_b200_contact;

typedef struct c2a_b200_model c2a_b200_model; /* device-resident model, opaque */
typedef struct c2a_b200_host_bvh c2a_b200_host_bvh; /* host-resident built hierarchy, opaque */

/* Host-side BVH build: what C2A_Model::BeginModel / AddTri / EndModel do for the CCD path
 * (C2A/src/C2A_PQP.cpp:82-121,228-290,331-417 -> C2A_BuildModel C2A/src/C2A_Build.cpp:546-574).
 * tris9: [n_tris][9] = p1,p2,p3 in AddTri order.  The tree is bit-identical to the reference's.
 * c2a_b200_bvh_view fills `view` with pointers into the built hierarchy (valid until
 * c2a_b200_bvh_free); tri_ids [n_tris] maps the builder's permuted triangle order back to the AddTri
 * index (PQP Tri::id). */
int c2a_b200_bvh_build(const double *tris9, int32_t n_tris, c2a_b200_host_bvh **out);
/* same, keeping the vertex indices of each triangle (AddTri's i1,i2,i3) for contact-feature labels */
int c2a_b200_bvh_build_indexed(const double *tris9, const int32_t *vidx3, int32_t n_tris, c2a_b200_host_bvh **out);
int c2a_b200_bvh_view(const c2a_b200_host_bvh *h, struct c2a_b200_bvh *view, const int32_t **tri_ids,
                      int32_t *depth);
int c2a_b200_bvh_free(c2a_b200_host_bvh *h);

/* Per-query outputs, structure of arrays; any pointer may be NULL to skip that output.
 * Host pointers for c2a_b200_solve_batch, device pointers for c2a_b200_solve_batch_device. */
typedef struct c2a_b200_results
{
  int32_t *status;        /* [n]     C2A_B200_QUERY_*                                                  */
  int32_t *collisionfree; /* [n]     C2A_TimeOfContactResult::collisionfree (the verdict, quirk Q1)    */
  int32_t *num_ca;        /* [n]     ::numCA == number_of_iteration of C2A_Solve                        */
  int32_t *num_bv_tests;  /* [n]     ::num_bv_tests                                                    */
  int32_t *num_tri_tests; /* [n]     ::num_tri_tests                                                   */
  double *toc;            /* [n]     ::toc == time_of_contact of C2A_Solve                              */
  double *distance;       /* [n]     ::distance                                                        */
  double *mint;           /* [n]     ::mint of the last CA step                                        */
  double *p1p2;           /* [n][6]  ::p1, ::p2 (closest points, model-1 frame)                        */
  double *pose_toc;       /* [n][24] trans0, trans1 of C2A_Solve as R(9)+T(3) each; only written when
                                     collisionfree == 0, like the reference                            */
  /* contact pass of C2A_Solve (C2A_QueryContact at the TOC pose with threshold 2*distance + 0.001,
   * C2A/src/C2A.cpp:2431-2434); run only when num_contact or contacts is non-NULL */
  int32_t *num_contact;   /* [n]     ::num_contact == number_of_contact of C2A_Solve (0 for free queries) */
  c2a_b200_contact *contacts; /* [n][max_contacts] in the traversal's visiting order (the reference's
                                     std::list is this order reversed: it push_front()s); the first
                                     min(num_contact, max_contacts) entries of each row are written */
  int32_t max_contacts;
  int32_t *last_tri;      /* [n][2]  triangle indices the traversal leaves in o1->last_tri / o2->last_tri
                                     (C2A/src/C2A.cpp:1175-1176): the last leaf that improved the distance;
                                     -1, -1 if none did.  The demo feeds them back as the next call's seeds
                                     (CCDDemo/mainTorusknot.cpp:314-315). */
} c2a_b200_results;

int c2a_b200_device_count(int32_t *count);

/* Upload a built model to `device`.  Replaces nothing in the reference (its models live in host
 * memory, C2A/C2A_Internal.h:32-81); the input is what C2A_Model::EndModel() produced
 * (C2A/src/C2A_PQP.cpp:331-417). */
int c2a_b200_model_upload(const c2a_b200_bvh *bvh, int32_t device, c2a_b200_model **out);
int c2a_b200_model_free(c2a_b200_model *m);
int c2a_b200_model_info(const c2a_b200_model *m, int32_t *device, int32_t *n_nodes, int32_t *n_tris,
                        int32_t *depth);

/* Batched C2A_Solve (C2A/src/C2A.cpp:2315-2444; its contact pass runs when num_contact / contacts is set) over n independent
 * queries on the models' device.  poses: [n][48] = trans00, trans01, trans10, trans11, each R (9,
 * row-major) + T (3) (the four Transform* of C2A_Solve).  seed_a / seed_b: [n] triangle indices
 * standing in for res->last_triA / last_triB (C2A/src/C2A.cpp:1816-1817; NULL = triangle 0, the
 * state after EndModel, C2A/src/C2A_PQP.cpp:401); every seed must be a triangle index of its model: the host-buffer
 * entries return C2A_B200_ERR_ARG otherwise, the device-pointer entry cannot inspect its seeds and treats an
 * out-of-range one as triangle 0.  tol_d / tol_t: C2A_Solve hard-codes 1e-4 for
 * both (C2A/src/C2A.cpp:2384-2385); C2A_QueryTimeOfContact takes them as arguments.
 * Queries whose two angular speeds are both < 1e-8 take the reference's translation-only branch
 * (C2A/src/C2A.cpp:2391-2395, :1362-1521; CInterpMotion::m_toc_delta = tol_d as C2A_Solve sets it): toc = the
 * step bound of its single traversal, collisionfree = (toc >= 1), num_ca = 0, last_tri = res->last_triA/B.
 * Host buffers; host<->device copies are inside the call. */
int c2a_b200_solve_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses,
                         const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d, double tol_t,
                         const c2a_b200_results *out);

/* The same batch sharded over n_devices GPUs of one box (no counterpart in the reference, whose
 * C2A_Solve is single-threaded: global b_TanslationCCD, C2A/src/C2A.cpp:33).  a[d], b[d]: replicas of the two models
 * uploaded to n_devices distinct devices.  Queries are independent, so there is no collective: the host computes the
 * motion constants and the claim order once, device d solves the queries order[d], order[d + n_devices], ... (the
 * cost-sorted order interleaved, so that long-running queries spread over the devices) on its own host thread and
 * stream, and the "gather" is each device's D2H followed by a scatter into the caller's arrays.  Host buffers;
 * results are identical to c2a_b200_solve_batch's for any n_devices.  The contact pass is not available here. */
int c2a_b200_solve_batch_multi(const c2a_b200_model *const *a, const c2a_b200_model *const *b, int32_t n_devices,
                               const double *poses, const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d,
                               double tol_t, const c2a_b200_results *out);

/* Batched C2A_QueryContact (C2A/C2A.h:284-289, C2A/src/C2A.cpp:1937-1966): all triangle pairs within
 * threshold[i] at the poses poses24[i] = pose of A, pose of B (R(9)+T(3) each).  Host buffers. */
int c2a_b200_contacts_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24,
                            const double *threshold, int64_t n, int32_t max_contacts, int32_t *num_contact,
                            c2a_b200_contact *contacts);

/* Batched C2A_Distance (C2A/C2A.h:256-261, C2A/src/C2A_PQP.cpp:970-1056): minimum distance between the two models
 * at the static poses poses24[i] = pose of A, pose of B (R(9)+T(3) each), with the depth-first routine the reference
 * takes for qsize <= 2 (C2ADistanceRecurse, :481-614; the result is visiting-order dependent exactly like the CCD
 * traversal, so a priority-queue run -- qsize > 2 -- may report another pair within the same error bounds).
 * seed_a / seed_b: [n] the models' last_tri going in (NULL = triangle 0); tri_pair [n][2]: the closest triangle pair
 * in builder order = last_tri coming out; p1p2 [n][6]: the closest points, each in its own model's frame
 * (PQP_DistanceResult::p1, p2).  rel_err / abs_err as PQP: a pair is skipped only if BOTH bounds allow it.
 * Host buffers; not part of the CCD hot path (it reuses its device functions). */
int c2a_b200_distance_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, const int32_t *seed_a,
                            const int32_t *seed_b, int64_t n, double rel_err, double abs_err, double *distance, double *p1p2,
                            int32_t *tri_pair, int32_t *num_bv_tests, int32_t *num_tri_tests);

/* The same query with the routine C2A_Distance takes for qsize > 2 (C2ADistanceQueueRecurse, C2A/src/C2A_PQP.cpp:624-787):
 * best-first over a bounded queue of pending node pairs (PQP's BVTQ, not in the reference's tree), recursing with a fresh
 * queue when it is full.  With both error bounds zero the distance equals the depth-first routine's; the reported pair,
 * points and counters are those of this visiting order.  Among equally distant pending pairs the one queued first is
 * taken first -- the behaviour of the queue stand-in the reference is compiled with here (oracle/pqp_shim/BVTQ.h); PQP's
 * own heap may break such ties differently.  qsize <= 2 runs the depth-first routine, like the reference (:1031). */
int c2a_b200_distance_queue_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24,
                                  const int32_t *seed_a, const int32_t *seed_b, int64_t n, double rel_err, double abs_err,
                                  int32_t qsize, double *distance, double *p1p2, int32_t *tri_pair, int32_t *num_bv_tests,
                                  int32_t *num_tri_tests);

/* Batched C2A_Collide, PQP_CollideResult overload (C2A/C2A.h:249-253, C2A/src/C2A_PQP.cpp:798-968): the pairs of
 * intersecting triangles of the two models at the static poses poses24[i].  flag: 1 = C2A_ALL_CONTACTS, 2 =
 * C2A_FIRST_CONTACT (C2A/C2A.h:246-247).  num_pairs [n] = PQP_CollideResult::NumPairs(); pairs [n][max_pairs][2]: the
 * first min(num_pairs, max_pairs) pairs of each query in the order the reference's traversal reports them, as
 * BUILDER-ORDER triangle indices (tri_ids of c2a_b200_bvh_view maps them to the reference's Tri::id = AddTri index);
 * unused entries are -1.  num_bv_tests / num_tri_tests as PQP_CollideResult (may be NULL; pairs may be NULL with
 * max_pairs 0).  Both models must have been uploaded with obb_d / obb_To.
 * The box and triangle overlap tests are PQP's obb_disjoint / TriContact (not in the reference's tree): restated from
 * the published separating-axis tests; parity is against the reference compiled with this repo's PQP stand-in
 * (oracle/pqp_shim), see DESIGN.md section 8.  Host buffers; not part of the CCD hot path. */
int c2a_b200_collide_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24, int64_t n, int32_t flag,
                           int32_t max_pairs, int32_t *num_pairs, int32_t *pairs, int32_t *num_bv_tests,
                           int32_t *num_tri_tests);

/* Batched C2A_Collide, C2A_DistanceResult overload (C2A/C2A.h:262-267, C2A/src/C2A_PQP.cpp:1060-1280): the walk of
 * c2a_b200_distance_batch restricted to node pairs whose boxes overlap (the result is the seed pair's distance when
 * the root boxes are apart).  Arguments and outputs as c2a_b200_distance_batch. */
int c2a_b200_collide_distance_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *poses24,
                                    const int32_t *seed_a, const int32_t *seed_b, int64_t n, double rel_err, double abs_err,
                                    double *distance, double *p1p2, int32_t *tri_pair, int32_t *num_bv_tests,
                                    int32_t *num_tri_tests);

/* Host half of the motion model: what constructing the two CInterpMotion_Linear objects does in
 * C2A_Solve (C2A/src/C2A.cpp:2378-2379 -> C2A/src/InterpMotion.cpp:148-168, 486-491, 228-270).
 * poses [n][48] -> motions [n][C2A_B200_MOTION_DOUBLES]: per object R0(9) T0(3) cv(3) axis(3) angVel
 * qs(4) m_toc_delta (0 = the batch's tol_d; only object 1's is read, by the translation-only branch).  It stays on the host because LinearAngularVelocity calls acos(): the reference's
 * constants are whatever the host libm returns, and the device consumes exactly those.
 * n_threads <= 0: all host cores. */
#define C2A_B200_MOTION_DOUBLES 48
int c2a_b200_motions_from_poses(const double *poses, int64_t n, double *motions, int32_t n_threads);

/* Batched C2A_QueryTimeOfContact (C2A/C2A.h:274-281) from ready motion records (host buffers): what
 * c2a_b200_solve_batch does after its motion set-up.  Used by the C++ shim of C2A_QueryTimeOfContact,
 * whose CInterpMotion arguments already hold cv / m_axis / m_angVel. */
int c2a_b200_solve_batch_motions(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions,
                                 const int32_t *seed_a, const int32_t *seed_b, int64_t n, double tol_d, double tol_t,
                                 const c2a_b200_results *out);

/* Batched C2A_TimeOfContactStep (C2A/src/C2A.cpp:1778-1931; note its argument order: tolerance_t, then
 * tolerance_d): ONE conservative-advancement iteration per query.  step_in [n][28] = the current poses
 * R1(9) T1(3) R2(9) T2(3), then res->numCA, res->mint (of the previous step) and res->UpboundTOC as the
 * step reads them, then a pad.  Writes distance, mint, p1p2, num_bv_tests, num_tri_tests, status. */
#define C2A_B200_STEP_IN_DOUBLES 28
int c2a_b200_toc_step_batch(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions,
                            const double *step_in, const int32_t *seed_a, const int32_t *seed_b, int64_t n,
                            double tol_t, double tol_d, const c2a_b200_results *out);

/* Batched C2A_QueryTimeOfContact (+ the pose outputs of C2A_Solve) with everything already resident
 * on the models' device: motions_dev [n][C2A_B200_MOTION_DOUBLES] from c2a_b200_motions_from_poses,
 * seeds, the optional claim order and outputs device pointers.  Enqueued on `cuda_stream` (a cudaStream_t; NULL = default
 * stream) without synchronising. */
int c2a_b200_solve_batch_device(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions_dev,
                                const int32_t *seed_a_dev, const int32_t *seed_b_dev, const int32_t *order_dev,
                                int64_t n, double tol_d, double tol_t, const c2a_b200_results *out_dev,
                                void *cuda_stream);

/* Claim order for a batch (host): order[k] = the query the k-th free worker takes.  Queries expected to
 * need many conservative-advancement steps (small closing speed relative to their motion bound) come
 * first, so their sequential chains do not form the tail of the launch.  Purely a scheduling hint:
 * results are independent of it.  The host-buffer entries compute it internally; pass it (uploaded) as
 * order_dev to c2a_b200_solve_batch_device, or NULL for index order. */
int c2a_b200_schedule_order(const c2a_b200_model *a, const c2a_b200_model *b, const double *motions, int64_t n,
                            int32_t *order);

/* Heterogeneous batch: per-query model handles (the reference has no batch API, its
 * C2A_Solve takes the two C2A_Model* per call, C2A/C2A.h:23-35).  models: [n_models] handles on ONE device;
 * model_a / model_b: [n] indices into it.  Queries are grouped by (model_a, model_b); each group is one launch of
 * the batched kernel over its own queries, results land at the query's index.  Otherwise as c2a_b200_solve_batch
 * (host buffers; the contact pass is not available here). */
int c2a_b200_solve_pairs(const c2a_b200_model *const *models, int32_t n_models, const int32_t *model_a,
                         const int32_t *model_b, const double *poses, const int32_t *seed_a, const int32_t *seed_b,
                         int64_t n, double tol_d, double tol_t, const c2a_b200_results *out);

/* Swept-sphere broadphase for a scene of moving instances (NOT in the reference: config 4 needs a candidate pair
 * list).  Instance i's model-frame origin moves on a straight line from c0[i] to c1[i] -- which is what
 * CInterpMotion_Linear does to it (C2A/src/InterpMotion.cpp:516-568) -- and all its vertices stay within
 * radius[i] of that origin (use max |vertex|, not C2A_Model::radius).  Reports every pair
 * i < j whose spheres come within `margin` of each other for some t in [0, 1]: conservative for the CCD query.
 * c0, c1: [n][3]; pairs: [max_pairs][2], unordered; *n_pairs = pairs found (if > max_pairs only max_pairs were
 * written).  Host buffers, brute force on `device` (n^2/2 sphere tests). */
int c2a_b200_broadphase(const double *c0, const double *c1, const double *radius, int32_t n, double margin, int32_t device,
                        int32_t *pairs, int64_t max_pairs, int64_t *n_pairs);

/* Number of kernel launches issued by this library on the calling process so far. */
int64_t c2a_b200_launch_count(void);

const char *c2a_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
