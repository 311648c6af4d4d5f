// c2a_b200 drop-in C++ API: the reference's public entry points and types for the continuous
// collision detection path, backed by the B200 CUDA library (include/c2a_b200.h).
//
// A program written against the reference (EwhaGlab/C2A) that includes "C2A/C2A.h",
// "C2A/LinearMath.h", "C2A/InterpMotion.h" and calls
//     C2A_Model::BeginModel / AddTri / EndModel          C2A/C2A_Internal.h:32-81, C2A/src/C2A_PQP.cpp:82-417
//     C2A_Solve                                          C2A/C2A.h:23-35,    C2A/src/C2A.cpp:2315-2444
//     C2A_QueryTimeOfContact                             C2A/C2A.h:274-281,  C2A/src/C2A.cpp:1987-2146
//     C2A_TimeOfContactStep                              (not in the reference header) C2A/src/C2A.cpp:1778-1789
//     CInterpMotion_Linear                               C2A/InterpMotion.h:164-226
// compiles against this directory instead and links libc2a_b200.so.  Same names, argument meaning
// and result fields; the traversal runs on the GPU.  Added: C2A_SolveBatch.
//
// Deliberate differences (each documented in DESIGN.md):
//   * nothing prints to stdout (the reference prints from EndModel and from every C2A_Solve call);
//   * contact features: FeatureID entries the reference leaves uninitialised are -1 here;
//   * the translation-only branch (both angular speeds < 1e-8, C2A/src/C2A.cpp:2391-2395) is chosen per
//     query from the two motions' m_angVel (the reference reads a global that C2A_Solve sets); where the
//     reference reads uninitialised memory in that branch (see c2a_b200/csrc/c2a_translation.cuh) the
//     behaviour of its object code under any finite non-zero garbage is reproduced; res->p1/p2 stay zero;
//   * degenerate rotations do not exit(0) (C2A/LinearMath.h:768);
//   * C2A_QueryTimeOfContact / C2A_TimeOfContactStep called DIRECTLY (C2A_Solve always constructs fresh motions, for
//     which none of this is visible): the first CA step starts from the motions' start pose transform_s, where the
//     reference reads their current pose ->transform (C2A/src/C2A.cpp:2015-2018) -- the two differ only if the caller
//     integrate()d a motion before the query; and after a rotational query the motions are not left integrate()d at
//     the last lamda the loop tried (:2111-2112) -- objmotion1 is at the TOC pose after a hit (:2143), objmotion2 and
//     the motions of a collision-free query keep the pose they had, so a C2A_QueryContact issued right after a direct
//     C2A_QueryTimeOfContact must integrate() them itself (C2A_Solve does: it runs its contact pass at the TOC poses).
#ifndef C2A_B200_DROPIN_C2A_H
#define C2A_B200_DROPIN_C2A_H

#include <list>
#include <utility>
#include <vector>

typedef double PQP_REAL;
typedef double Real;

const int PQP_OK = 0;
const int PQP_ERR_MODEL_OUT_OF_MEMORY = -1;
const int PQP_ERR_OUT_OF_MEMORY = -2;
const int PQP_ERR_UNPROCESSED_MODEL = -3;
const int PQP_ERR_BUILD_OUT_OF_SEQUENCE = -4;
const int PQP_ERR_BUILD_EMPTY_MODEL = -5;

// ---- value types of the public API (C2A/LinearMath.h; SWIFT++ lineage in the reference) ------------
class Coord3D
{
public:
  Coord3D() { val[0] = val[1] = val[2] = 0.0; }
  Coord3D(Real x, Real y, Real z) { val[0] = x; val[1] = y; val[2] = z; }
  explicit Coord3D(const Real v[]) { val[0] = v[0]; val[1] = v[1]; val[2] = v[2]; }
  Real &X() { return val[0]; }
  Real &Y() { return val[1]; }
  Real &Z() { return val[2]; }
  Real X() const { return val[0]; }
  Real Y() const { return val[1]; }
  Real Z() const { return val[2]; }
  Real &operator[](int i) { return val[i]; }
  Real operator[](int i) const { return val[i]; }
  void Get_Value(Real v[]) const { v[0] = val[0]; v[1] = val[1]; v[2] = val[2]; }
  void Set_Value(const Real v[]) { val[0] = v[0]; val[1] = v[1]; val[2] = v[2]; }
  void Set_Value(Real x, Real y, Real z) { val[0] = x; val[1] = y; val[2] = z; }
  void Identity() { val[0] = val[1] = val[2] = 0.0; }
  Real Length_Sq() const { return val[0] * val[0] + val[1] * val[1] + val[2] * val[2]; }
  Real val[3];
};

class Quaternion  // (x, y, z, w)
{
public:
  Quaternion() { val[0] = val[1] = val[2] = 0.0; val[3] = 1.0; }
  Quaternion(Real x, Real y, Real z, Real w) { val[0] = x; val[1] = y; val[2] = z; val[3] = w; }
  Real &X() { return val[0]; }
  Real &Y() { return val[1]; }
  Real &Z() { return val[2]; }
  Real &W() { return val[3]; }
  Real X() const { return val[0]; }
  Real Y() const { return val[1]; }
  Real Z() const { return val[2]; }
  Real W() const { return val[3]; }
  Real &operator[](int i) { return val[i]; }
  Real operator[](int i) const { return val[i]; }
  void Set_Value(Real x, Real y, Real z, Real w) { val[0] = x; val[1] = y; val[2] = z; val[3] = w; }
  Real val[4];
};
Quaternion operator%(const Quaternion &a, const Quaternion &b);  // Hamilton product, LinearMath.h:1018-1025

class Matrix3x3  // row-major
{
public:
  Matrix3x3() { Identity(); }
  Real *operator[](int i) { return &val[3 * i]; }
  const Real *operator[](int i) const { return &val[3 * i]; }
  void Get_Value(Real v[3][3]) const;
  void Set_Value(const Real v[3][3]);
  void Set_Value(const Real v[]);
  void Set_Value(const Quaternion &q);  // LinearMath.h:809-831
  Quaternion Quaternion_() const;      // LinearMath.h:759-793
  void Identity() { for (int i = 0; i < 9; i++) val[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  Real val[9];
};

class Transform
{
public:
  Transform() {}
  const Coord3D &Translation() const { return T; }
  Coord3D &Translation() { return T; }
  const Matrix3x3 &Rotation() const { return R; }
  Matrix3x3 &Rotation() { return R; }
  Quaternion Quaternion_() const { return R.Quaternion_(); }
  void Set_Value(const Real v[]);                 // row-major 3x4, 4th column = translation (LinearMath.h:917-925)
  void Set_Value(const Real r[], const Real t[]) { R.Set_Value(r); T.Set_Value(t); }
  void Set_Value(const Matrix3x3 &r, const Coord3D &t) { R = r; T = t; }
  void Set_Rotation(const Matrix3x3 &r) { R = r; }
  void Set_Rotation(const Quaternion &q) { R.Set_Value(q); }
  void Set_Translation(const Coord3D &t) { T = t; }
  void Identity() { R.Identity(); T.Identity(); }
  Matrix3x3 R;
  Coord3D T;
};

// ---- model (PQP Tri / C2A_Tri / C2A_Model) -----------------------------------------------------------
struct Tri
{
  PQP_REAL p1[3], p2[3], p3[3];
  int id;
};
struct C2A_Tri : public Tri
{
  int *Index() { return index_; }
  int index_[3];
};

struct c2a_b200_model;
struct c2a_b200_host_bvh;

class C2A_Model
{
public:
  C2A_Model();
  ~C2A_Model();
  int BeginModel(int n = 8);
  int AddTri(const PQP_REAL *p1, const PQP_REAL *p2, const PQP_REAL *p3, int id, int i1, int i2, int i3);
  int AddTri(const PQP_REAL *p1, const PQP_REAL *p2, const PQP_REAL *p3, int id);
  int EndModel();  // builds the RSS BVH on the host (bit-identical tree) and uploads it to the GPU
  int MemUsage(int msg);
  C2A_Tri *GetTriangle(int idx) { return &tris[idx]; }

  int build_state;
  C2A_Tri *tris;   // after EndModel: in the builder's (permuted) order, like the reference
  int num_tris;
  int num_bvs;
  Tri *last_tri;   // closest triangle of the last query; EndModel sets it to tris (C2A_PQP.cpp:401)

  // c2a_b200 additions
  int device;                   // GPU the model is uploaded to (set before EndModel; default 0)
  c2a_b200_model *gpu;          // device-resident hierarchy
  c2a_b200_host_bvh *host_bvh;  // host copy of the flattened hierarchy
  // Replica of the built hierarchy on another GPU of the box (for C2A_SolveBatchMulti); idempotent per device.
  // Returns PQP_OK, PQP_ERR_UNPROCESSED_MODEL before EndModel(), PQP_ERR_MODEL_OUT_OF_MEMORY if the upload fails.
  int ReplicateTo(int other_device);
  c2a_b200_model *OnDevice(int d) const;  // the replica on device d (the model itself for d == device), or 0
private:
  std::vector<std::pair<int, c2a_b200_model *> > replicas_;
  std::vector<C2A_Tri> storage_;
  C2A_Model(const C2A_Model &);
  C2A_Model &operator=(const C2A_Model &);
};

// ---- motion (C2A/InterpMotion.h) -----------------------------------------------------------------------
enum GMP_INTERP_MODE { GMP_IM_EULER = 0, GMP_IM_LINEAR = 1, GMP_IM_SCREW = 2, GMP_IM_SLERP = 3 };

class CInterpMotion
{
public:
  CInterpMotion(GMP_INTERP_MODE itpMode, const PQP_REAL R0[3][3], const PQP_REAL T0[3], const PQP_REAL R1[3][3],
                const PQP_REAL T1[3]);
  CInterpMotion();
  virtual ~CInterpMotion();
  virtual void velocity(void) {}
  virtual bool integrate(const double dt, PQP_REAL qua[7]) = 0;
  bool integrate(const double dt, PQP_REAL R[3][3], PQP_REAL T[3]);
  virtual double computeTOC(PQP_REAL d, PQP_REAL r1, PQP_REAL S[3]) { return 0; }
  void LinearAngularVelocity(Coord3D &axis, Real &angVel);
  Quaternion DeltaRt(Real t);
  Quaternion AbsoluteRt(Real t);

  Real m_toc_delta;
  GMP_INTERP_MODE m_itpMode;
  Transform transform;    // current pose (mutated by integrate)
  Transform transform_s;  // start
  Transform transform_t;  // end
  Quaternion quaternion_s, quaternion_t;
  Coord3D cv;      // linear velocity of the origin
  Coord3D m_axis;  // unit rotation axis
  Real m_angVel;   // angular speed about m_axis
};

class CInterpMotion_Linear : public CInterpMotion
{
public:
  CInterpMotion_Linear(const PQP_REAL R0[3][3], const PQP_REAL T0[3], const PQP_REAL R1[3][3], const PQP_REAL T1[3]);
  virtual ~CInterpMotion_Linear();
  virtual void velocity(void);
  virtual bool integrate(const double dt, PQP_REAL qua[7]);
  using CInterpMotion::integrate;
  virtual double computeTOC(PQP_REAL d, PQP_REAL r1, PQP_REAL S[3]);            // InterpMotion.cpp:746-828
  double computeTOC_MotionBound(PQP_REAL T[3], PQP_REAL d, PQP_REAL angularRadius, PQP_REAL N[3]);  // :831-881 (takes the
                                                                                // radius instead of a C2A_BV*)
};

// ---- query API (C2A/C2A.h) ---------------------------------------------------------------------------
enum C2A_Result { OK, TOCFound, CollisionFound, CollisionFree, CollisionNotFound };

struct ContactF
{
  int FeatureType_A, FeatureType_B;
  int FeatureID_A[3], FeatureID_B[3];
  int TriangleID_A, TriangleID_B;
  PQP_REAL P_A[3], P_B[3];
  PQP_REAL Distance;
};
typedef std::list<ContactF> ContactFList;

struct C2A_TimeOfContactResult
{
  int num_bv_tests;
  int num_tri_tests;
  double query_time_secs;
  ContactFList cont_l;
  PQP_REAL R[3][3];
  PQP_REAL T[3];
  PQP_REAL rel_err;
  PQP_REAL abs_err;
  PQP_REAL UpboundTOC;
  PQP_REAL distance;
  PQP_REAL p1[3];
  PQP_REAL p2[3];
  int qsize;
  PQP_REAL mint;
  PQP_REAL toc;        // 0 both for "collision free" and for toc >= 1 - tolerance_t (reference quirk Q1)
  bool collisionfree;  // the verdict
  int numCA;
  PQP_REAL R_toc[3][3], T_toc[3];
  int num_contact;
  Tri *last_triA;      // seed triangles of the next query (the reference never updates them itself)
  Tri *last_triB;

  C2A_TimeOfContactResult() : cont_l(), numCA(0), last_triA(0), last_triB(0) {}
  int NumBVTests() { return num_bv_tests; }
  int NumTriTests() { return num_tri_tests; }
  double QueryTimeSecs() { return query_time_secs; }
  PQP_REAL Distance() { return distance; }
  const PQP_REAL *P1() { return p1; }
  const PQP_REAL *P2() { return p2; }
};

// PQP_DistanceResult + the closest triangle pair (C2A/C2A_Internal.h:89-93)
struct C2A_DistanceResult
{
  int num_bv_tests, num_tri_tests;
  double query_time_secs;
  PQP_REAL R[3][3], T[3];   // model 2 -> model 1
  PQP_REAL rel_err, abs_err;
  PQP_REAL distance;
  PQP_REAL p1[3], p2[3];    // closest points, each in its own model's frame
  int qsize;
  int t1, t2;               // Tri::id of the closest pair
  int NumBVTests() { return num_bv_tests; }
  int NumTriTests() { return num_tri_tests; }
  double QueryTimeSecs() { return query_time_secs; }
  PQP_REAL Distance() { return distance; }
  const PQP_REAL *P1() { return p1; }
  const PQP_REAL *P2() { return p2; }
};

// C2A/C2A.h:256-261, C2A/src/C2A_PQP.cpp:970-1056: the depth-first routine for qsize <= 2, the priority-queue one
// (C2ADistanceQueueRecurse, :624-787) above, like the reference; equally distant pending pairs leave the queue in the
// order they entered it (PQP's own queue is not in the reference's tree and may break such ties differently).
// Reads and updates o1->last_tri / o2->last_tri like the reference.
int C2A_Distance(C2A_DistanceResult *result, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                 PQP_REAL T2[3], C2A_Model *o2, PQP_REAL rel_err, PQP_REAL abs_err, int qsize = 2);

// PQP's collision result (PQP.h; the reference takes it from the un-vendored PQP): the pairs of intersecting triangles
struct CollisionPair { int id1, id2; };
struct PQP_CollideResult
{
  int num_bv_tests, num_tri_tests;
  double query_time_secs;
  PQP_REAL R[3][3], T[3];   // model 2 -> model 1
  int num_pairs_alloced, num_pairs;
  CollisionPair *pairs;
  PQP_CollideResult() : num_bv_tests(0), num_tri_tests(0), query_time_secs(0), num_pairs_alloced(0), num_pairs(0), pairs(0) {}
  ~PQP_CollideResult() { FreePairsList(); }
  PQP_CollideResult(const PQP_CollideResult &) = delete;
  PQP_CollideResult &operator=(const PQP_CollideResult &) = delete;
  void SizeTo(int n);
  void Add(int i1, int i2);
  void FreePairsList() { delete[] pairs; pairs = 0; num_pairs = num_pairs_alloced = 0; }
  int NumBVTests() { return num_bv_tests; }
  int NumTriTests() { return num_tri_tests; }
  double QueryTimeSecs() { return query_time_secs; }
  int Colliding() { return num_pairs > 0; }
  int NumPairs() { return num_pairs; }
  int Id1(int k) { return pairs[k].id1; }
  int Id2(int k) { return pairs[k].id2; }
};

const int C2A_ALL_CONTACTS = 1;   // find all pairwise intersecting triangles (C2A/C2A.h:246)
const int C2A_FIRST_CONTACT = 2;  // report the first intersecting pair found (:247)

// C2A/C2A.h:249-253, C2A/src/C2A_PQP.cpp:798-968: the intersecting triangle pairs (Tri::id) in the order the reference's
// traversal reports them.  The box and triangle overlap tests are PQP's, restated.
int C2A_Collide(PQP_CollideResult *result, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                PQP_REAL T2[3], C2A_Model *o2, int flag = C2A_ALL_CONTACTS);

// C2A/C2A.h:262-267, C2A/src/C2A_PQP.cpp:1060-1280: C2A_Distance's walk restricted to node pairs whose boxes overlap.
// Reads and updates o1->last_tri / o2->last_tri like the reference.
int C2A_Collide(C2A_DistanceResult *result, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                PQP_REAL T2[3], C2A_Model *o2, PQP_REAL rel_err, PQP_REAL abs_err, int qsize = 2);

C2A_Result C2A_Solve(Transform *trans00, Transform *trans01, C2A_Model *obj1_tested, Transform *trans10,
                     Transform *trans11, C2A_Model *obj2_tested, Transform &trans0, Transform &trans1,
                     PQP_REAL &time_of_contact, int &number_of_iteration, int &number_of_contact, PQP_REAL th_ca,
                     C2A_TimeOfContactResult &dres);

PQP_REAL C2A_QueryTimeOfContact(CInterpMotion *objmotion1, CInterpMotion *objmotion2, C2A_TimeOfContactResult *res,
                                C2A_Model *o1, C2A_Model *o2, PQP_REAL tolerance_d, PQP_REAL tolerance_t, int qsize = 2);

PQP_REAL C2A_QueryContact(CInterpMotion *objmotion1, CInterpMotion *objmotion2, C2A_TimeOfContactResult *res, C2A_Model *o1,
                          C2A_Model *o2, double threshold);

// C2A/C2A.h:292, C2A/src/C2A.cpp:1969-1985: the contact pass at explicit poses (clears res->cont_l first)
PQP_REAL C2A_QueryContactOnly(C2A_TimeOfContactResult *res, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                              PQP_REAL T2[3], C2A_Model *o2, double threshold);

int C2A_TimeOfContactStep(CInterpMotion *objmotion1, CInterpMotion *objmotion2, C2A_TimeOfContactResult *res,
                          PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3], PQP_REAL T2[3],
                          C2A_Model *o2, PQP_REAL tolerance_t, PQP_REAL tolerance_d);

// New: n independent C2A_Solve queries of the same object pair in one GPU launch.  trans00..trans11 are
// arrays of n Transforms; seed_tri_a/b (optional, n entries) are triangle indices standing in for
// dres.last_triA/B (NULL: triangle 0).  Outputs are arrays of n (any may be NULL): collisionfree,
// time_of_contact, distance, number_of_iteration, trans0/trans1 (poses at TOC, written for hits only).
// Returns PQP_OK or a negative c2a_b200 error code.
int C2A_SolveBatch(int n, const Transform *trans00, const Transform *trans01, C2A_Model *obj1_tested,
                   const Transform *trans10, const Transform *trans11, C2A_Model *obj2_tested,
                   const int *seed_tri_a, const int *seed_tri_b, bool *collisionfree, PQP_REAL *time_of_contact,
                   PQP_REAL *distance, int *number_of_iteration, Transform *trans0, Transform *trans1);

// The same batch sharded over several GPUs of the box (c2a_b200_solve_batch_multi): devices[0..n_devices) are distinct
// CUDA devices on each of which both models have a replica (the device they were built for, or C2A_Model::ReplicateTo).
// Results are identical to C2A_SolveBatch's.
int C2A_SolveBatchMulti(const int *devices, int n_devices, int n, const Transform *trans00, const Transform *trans01,
                        C2A_Model *obj1_tested, const Transform *trans10, const Transform *trans11, C2A_Model *obj2_tested,
                        const int *seed_tri_a, const int *seed_tri_b, bool *collisionfree, PQP_REAL *time_of_contact,
                        PQP_REAL *distance, int *number_of_iteration, Transform *trans0, Transform *trans1);

#endif
