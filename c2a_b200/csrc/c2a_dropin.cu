// Host-side mirror of the reference's C++ interface for the CCD path (include/C2A/C2A.h): same names,
// argument meaning and result fields as /root/reference/C2A/C2A.h, C2A/InterpMotion.h,
// C2A/C2A_Internal.h, C2A/LinearMath.h; the traversal itself is delegated to the C ABI
// (include/c2a_b200.h).  Host arithmetic keeps the reference's operation order (compile with
// -ffp-contract=off), and uses the host libm for sin/cos/acos exactly like the reference.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../../include/C2A/C2A.h"
#include "../../include/c2a_b200.h"
#include "c2a_motion.cuh"  // quat_from_matrix / quat_mul (host + device)

enum { C2A_BUILD_STATE_EMPTY = 0, C2A_BUILD_STATE_BEGUN = 1, C2A_BUILD_STATE_PROCESSED = 2 };

// ---- LinearMath subset -----------------------------------------------------------------------------
Quaternion operator%(const Quaternion &a, const Quaternion &b)
{
  Quaternion r;
  c2a::quat_mul(r.val, a.val, b.val);
  return r;
}
void Matrix3x3::Get_Value(Real v[3][3]) const
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) v[i][j] = val[3 * i + j];
}
void Matrix3x3::Set_Value(const Real v[3][3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) val[3 * i + j] = v[i][j];
}
void Matrix3x3::Set_Value(const Real v[])
{
  for (int i = 0; i < 9; i++) val[i] = v[i];
}
// Matrix3x3::Set_Value(Quaternion), C2A/LinearMath.h:809-831
void Matrix3x3::Set_Value(const Quaternion &q)
{
  const Real d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const Real s = 2.0 / d;
  const Real xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
  const Real wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  const Real xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
  const Real yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  val[0] = 1.0 - (yy + zz); val[1] = xy - wz; val[2] = xz + wy;
  val[3] = xy + wz; val[4] = 1.0 - (xx + zz); val[5] = yz - wx;
  val[6] = xz - wy; val[7] = yz + wx; val[8] = 1.0 - (xx + yy);
}
Quaternion Matrix3x3::Quaternion_() const
{
  Quaternion q;
  c2a::quat_from_matrix(q.val, val);
  return q;
}
void Transform::Set_Value(const Real v[])
{
  R.val[0] = v[0]; R.val[1] = v[1]; R.val[2] = v[2];
  R.val[3] = v[4]; R.val[4] = v[5]; R.val[5] = v[6];
  R.val[6] = v[8]; R.val[7] = v[9]; R.val[8] = v[10];
  T.val[0] = v[3]; T.val[1] = v[7]; T.val[2] = v[11];
}

// ---- C2A_Model ---------------------------------------------------------------------------------------
C2A_Model::C2A_Model()
    : build_state(C2A_BUILD_STATE_EMPTY), tris(0), num_tris(0), num_bvs(0), last_tri(0), device(0), gpu(0), host_bvh(0)
{
}
C2A_Model::~C2A_Model()
{
  for (size_t i = 0; i < replicas_.size(); i++) c2a_b200_model_free(replicas_[i].second);
  if (gpu) c2a_b200_model_free(gpu);
  if (host_bvh) c2a_b200_bvh_free(host_bvh);
}
int C2A_Model::ReplicateTo(int other_device)
{
  if (build_state != C2A_BUILD_STATE_PROCESSED || !host_bvh) return PQP_ERR_UNPROCESSED_MODEL;
  if (OnDevice(other_device)) return PQP_OK;
  c2a_b200_bvh view;
  c2a_b200_bvh_view(host_bvh, &view, 0, 0);
  c2a_b200_model *m = 0;
  if (c2a_b200_model_upload(&view, other_device, &m))
  {
    fprintf(stderr, "c2a_b200: model upload failed: %s\n", c2a_b200_last_error());
    return PQP_ERR_MODEL_OUT_OF_MEMORY;
  }
  replicas_.push_back(std::make_pair(other_device, m));
  return PQP_OK;
}
c2a_b200_model *C2A_Model::OnDevice(int d) const
{
  if (d == device) return gpu;
  for (size_t i = 0; i < replicas_.size(); i++)
    if (replicas_[i].first == d) return replicas_[i].second;
  return 0;
}
int C2A_Model::BeginModel(int n)
{
  const bool was_empty = build_state == C2A_BUILD_STATE_EMPTY;
  if (!was_empty)
  {
    for (size_t i = 0; i < replicas_.size(); i++) c2a_b200_model_free(replicas_[i].second);
    replicas_.clear();
    if (gpu) { c2a_b200_model_free(gpu); gpu = 0; }
    if (host_bvh) { c2a_b200_bvh_free(host_bvh); host_bvh = 0; }
    storage_.clear();
    num_tris = num_bvs = 0;
  }
  storage_.reserve(n > 0 ? n : 8);
  tris = 0; last_tri = 0;
  build_state = C2A_BUILD_STATE_BEGUN;
  if (!was_empty)
  {
    fprintf(stderr, "PQP Warning! Called BeginModel() on a PQP_Model that \nwas not empty. This model was cleared and "
                    "previous\ntriangle additions were lost.\n");
    return PQP_ERR_BUILD_OUT_OF_SEQUENCE;
  }
  return PQP_OK;
}
int C2A_Model::AddTri(const PQP_REAL *p1, const PQP_REAL *p2, const PQP_REAL *p3, int id, int i1, int i2, int i3)
{
  if (build_state == C2A_BUILD_STATE_EMPTY) BeginModel();
  else if (build_state == C2A_BUILD_STATE_PROCESSED)
  {
    fprintf(stderr, "PQP Warning! Called AddTri() on C2A_Model \nobject that was already ended. AddTri() was\nignored.  "
                    "Must do a BeginModel() to clear the\nmodel for addition of new triangles\n");
    return PQP_ERR_BUILD_OUT_OF_SEQUENCE;
  }
  C2A_Tri t;
  for (int k = 0; k < 3; k++) { t.p1[k] = p1[k]; t.p2[k] = p2[k]; t.p3[k] = p3[k]; }
  t.id = id;
  t.index_[0] = i1; t.index_[1] = i2; t.index_[2] = i3;
  storage_.push_back(t);
  num_tris = (int)storage_.size();
  tris = storage_.data();
  return PQP_OK;
}
int C2A_Model::AddTri(const PQP_REAL *p1, const PQP_REAL *p2, const PQP_REAL *p3, int id)
{
  return AddTri(p1, p2, p3, id, 0, 0, 0);
}
int C2A_Model::EndModel()
{
  if (build_state == C2A_BUILD_STATE_PROCESSED)
  {
    fprintf(stderr, "PQP Warning! Called EndModel() on C2A_Model \nobject that was already ended. EndModel() was\n"
                    "ignored.  Must do a BeginModel() to clear the\nmodel for addition of new triangles\n");
    return PQP_ERR_BUILD_OUT_OF_SEQUENCE;
  }
  if (num_tris == 0)
  {
    fprintf(stderr, "PQP Error! EndModel() called on model with no triangles\n");
    return PQP_ERR_BUILD_EMPTY_MODEL;
  }
  std::vector<double> t9((size_t)9 * num_tris);
  for (int i = 0; i < num_tris; i++)
    for (int k = 0; k < 3; k++) { t9[9 * (size_t)i + k] = storage_[i].p1[k]; t9[9 * (size_t)i + 3 + k] = storage_[i].p2[k]; t9[9 * (size_t)i + 6 + k] = storage_[i].p3[k]; }
  std::vector<int32_t> vi((size_t)3 * num_tris);
  for (int i = 0; i < num_tris; i++)
    for (int k = 0; k < 3; k++) vi[3 * (size_t)i + k] = storage_[i].index_[k];
  if (host_bvh) { c2a_b200_bvh_free(host_bvh); host_bvh = 0; }   // (left by an EndModel() that failed to upload)
  int rc = c2a_b200_bvh_build_indexed(t9.data(), vi.data(), num_tris, &host_bvh);
  if (rc) return PQP_ERR_MODEL_OUT_OF_MEMORY;
  c2a_b200_bvh view;
  const int32_t *ids = 0;
  c2a_b200_bvh_view(host_bvh, &view, &ids, 0);
  rc = c2a_b200_model_upload(&view, device, &gpu);
  if (rc)
  {
    // nothing of the model has changed yet: EndModel() can be called again (e.g. after freeing device memory)
    fprintf(stderr, "c2a_b200: model upload failed: %s\n", c2a_b200_last_error());
    c2a_b200_bvh_free(host_bvh); host_bvh = 0;
    return PQP_ERR_MODEL_OUT_OF_MEMORY;
  }
  // like the reference, the triangle array ends up in the builder's permuted order (Tri::id keeps the AddTri index)
  std::vector<C2A_Tri> perm(num_tris);
  for (int i = 0; i < num_tris; i++) perm[i] = storage_[ids[i]];
  storage_.swap(perm);
  tris = storage_.data();
  num_bvs = view.n_nodes;
  build_state = C2A_BUILD_STATE_PROCESSED;
  last_tri = tris;
  return PQP_OK;
}
int C2A_Model::MemUsage(int msg)
{
  const int total = (int)(sizeof(C2A_Tri) * num_tris + 224 * (size_t)num_bvs + sizeof(C2A_Model));
  if (msg) fprintf(stderr, "Total for model %p: %d bytes\n", (void *)this, total);
  return total;
}

// ---- CInterpMotion -----------------------------------------------------------------------------------
CInterpMotion::CInterpMotion() : m_toc_delta(0), m_itpMode(GMP_IM_LINEAR), m_angVel(0) {}
CInterpMotion::~CInterpMotion() {}
CInterpMotion::CInterpMotion(GMP_INTERP_MODE itpMode, const PQP_REAL R0[3][3], const PQP_REAL T0[3], const PQP_REAL R1[3][3],
                             const PQP_REAL T1[3])
    : m_toc_delta(0), m_itpMode(itpMode), m_angVel(0)
{
  Real r0[9], r1[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { r0[3 * i + j] = R0[i][j]; r1[3 * i + j] = R1[i][j]; }
  transform.Set_Value(r0, T0);
  transform_s.Set_Value(r0, T0);
  transform_t.Set_Value(r1, T1);
}
// C2A/src/InterpMotion.cpp:228-270
void CInterpMotion::LinearAngularVelocity(Coord3D &axis, Real &angVel)
{
  quaternion_s = transform_s.Rotation().Quaternion_();
  quaternion_t = transform_t.Rotation().Quaternion_();
  Quaternion s_q0(-quaternion_s.X(), -quaternion_s.Y(), -quaternion_s.Z(), quaternion_s.W());
  Quaternion s_temp = s_q0 % quaternion_t;
  const double s = 1 < s_temp.W() ? 1 : s_temp.W();
  const double sign = s < 0 ? -1 : 1;
  const double a = (fabs(s - 1) <= 1e-40 || fabs(s + 1) <= 1e-40) ? (2 * sign) : (sign * acos(2 * s * s - 1) / sqrt(1 - s * s));
  double tangent[3] = {a * s_temp.X(), a * s_temp.Y(), a * s_temp.Z()};
  axis.Set_Value(tangent);
  angVel = sqrt(tangent[0] * tangent[0] + tangent[1] * tangent[1] + tangent[2] * tangent[2]);
  const PQP_REAL len = axis.Length_Sq();
  if (len < 0.00000001f) axis.Set_Value(1.f, 0.f, 0.f);
  else
  {
    const Real inv = 1.0 / sqrt(len);
    axis.val[0] *= inv; axis.val[1] *= inv; axis.val[2] *= inv;
  }
}
Quaternion CInterpMotion::DeltaRt(Real t)
{
  const Real ang = 0.5f * m_angVel * t;
  const Real sn = sin(ang);
  return Quaternion(sn * m_axis.X(), sn * m_axis.Y(), sn * m_axis.Z(), cos(ang));
}
Quaternion CInterpMotion::AbsoluteRt(Real t)
{
  Quaternion d_rt = DeltaRt(t);
  Quaternion orn0 = transform_s.Quaternion_();
  return orn0 % d_rt;
}
bool CInterpMotion::integrate(const double dt, PQP_REAL R[3][3], PQP_REAL T[3])
{
  PQP_REAL qua[7];
  integrate(dt, qua);
  transform.Rotation().Get_Value(R);
  transform.Translation().Get_Value(T);
  return true;
}

CInterpMotion_Linear::CInterpMotion_Linear(const PQP_REAL R0[3][3], const PQP_REAL T0[3], const PQP_REAL R1[3][3],
                                           const PQP_REAL T1[3])
    : CInterpMotion(GMP_IM_LINEAR, R0, T0, R1, T1)
{
  velocity();
}
CInterpMotion_Linear::~CInterpMotion_Linear() {}
void CInterpMotion_Linear::velocity(void)
{
  for (int i = 0; i < 3; i++) cv[i] = transform_t.Translation()[i] - transform_s.Translation()[i];
  LinearAngularVelocity(m_axis, m_angVel);
}
// C2A/src/InterpMotion.cpp:516-568
bool CInterpMotion_Linear::integrate(const double dt_input, PQP_REAL qua[7])
{
  double dt = dt_input;
  if (dt > 1) dt = 1;
  for (int i = 0; i < 3; i++) transform.Translation()[i] = transform_s.Translation()[i] + dt * cv[i];
  Quaternion predictedOrn = AbsoluteRt(dt);
  transform.Set_Rotation(predictedOrn);
  qua[0] = predictedOrn.W(); qua[1] = predictedOrn.X(); qua[2] = predictedOrn.Y(); qua[3] = predictedOrn.Z();
  qua[4] = transform.Translation().X(); qua[5] = transform.Translation().Y(); qua[6] = transform.Translation().Z();
  return true;
}
static inline void normalize3(PQP_REAL v[3])
{
  const PQP_REAL d = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  v[0] *= d; v[1] *= d; v[2] *= d;
}
double CInterpMotion_Linear::computeTOC(PQP_REAL d, PQP_REAL r1, PQP_REAL S[3])
{
  PQP_REAL v_max, w_max;
  normalize3(S);
  if (m_angVel == 0)
  {
    w_max = 0;
    v_max = (cv[0] * S[0] + cv[1] * S[1] + cv[2] * S[2]);
    if (v_max < 0) v_max = 0;
  }
  else
  {
    PQP_REAL cwc[3] = {m_axis[0], m_axis[1], m_axis[2]}, cross[3];
    cwc[0] *= m_angVel; cwc[1] *= m_angVel; cwc[2] *= m_angVel;
    cross[0] = cwc[1] * S[2] - cwc[2] * S[1]; cross[1] = cwc[2] * S[0] - cwc[0] * S[2]; cross[2] = cwc[0] * S[1] - cwc[1] * S[0];
    w_max = r1 * sqrt(cross[0] * cross[0] + cross[1] * cross[1] + cross[2] * cross[2]);
    v_max = (cv[0] * S[0] + cv[1] * S[1] + cv[2] * S[2]);
    if (v_max < 0) v_max = 0;
  }
  PQP_REAL path_max = w_max + v_max;
  if (path_max == 0) path_max = 1e-30;
  return path_max;
}
double CInterpMotion_Linear::computeTOC_MotionBound(PQP_REAL T[3], PQP_REAL d, PQP_REAL angularRadius, PQP_REAL N[3])
{
  PQP_REAL cross[3];
  normalize3(N);
  cross[0] = m_axis[1] * N[2] - m_axis[2] * N[1]; cross[1] = m_axis[2] * N[0] - m_axis[0] * N[2]; cross[2] = m_axis[0] * N[1] - m_axis[1] * N[0];
  const PQP_REAL w_max = (angularRadius)*sqrt(cross[0] * cross[0] + cross[1] * cross[1] + cross[2] * cross[2]) * m_angVel;
  PQP_REAL v_max = (cv[0] * N[0] + cv[1] * N[1] + cv[2] * N[2]);
  if (v_max < 0) v_max = 0;
  PQP_REAL path_max = v_max + w_max;
  if (path_max <= 0) path_max = 1e-30;
  return path_max;
}

// ---- queries -----------------------------------------------------------------------------------------
namespace {
// motion record (c2a_motion.cuh): R0(9) T0(3) cv(3) axis(3) w qs(4) pad, from an already-built CInterpMotion
void record_from_motion(CInterpMotion *m, double *rec)
{
  for (int i = 0; i < 9; i++) rec[i] = m->transform_s.Rotation().val[i];
  for (int i = 0; i < 3; i++) { rec[9 + i] = m->transform_s.Translation()[i]; rec[12 + i] = m->cv[i]; rec[15 + i] = m->m_axis[i]; }
  rec[18] = m->m_angVel;
  const Quaternion qs = m->transform_s.Quaternion_();
  for (int i = 0; i < 4; i++) rec[19 + i] = qs[i];
  rec[23] = m->m_toc_delta;  // read by the translation-only branch (ConservD, InterpMotion.cpp:291-298); 0 = tolerance_d
}
int seed_index(C2A_Model *o, Tri *t)
{
  if (!t) return 0;
  const long idx = (C2A_Tri *)t - o->tris;
  return (idx >= 0 && idx < o->num_tris) ? (int)idx : 0;
}
void pose12(const Transform &t, double *p)
{
  for (int i = 0; i < 9; i++) p[i] = t.Rotation().val[i];
  for (int i = 0; i < 3; i++) p[9 + i] = t.Translation()[i];
}
}  // namespace

// C2A/src/C2A.cpp:1987-2146
PQP_REAL C2A_QueryTimeOfContact(CInterpMotion *objmotion1, CInterpMotion *objmotion2, C2A_TimeOfContactResult *res,
                                C2A_Model *o1, C2A_Model *o2, PQP_REAL tolerance_d, PQP_REAL tolerance_t, int qsize)
{
  (void)qsize;  // accepted and ignored, like the reference (C2A.cpp:1987-1995)
  res->num_bv_tests = 0; res->num_tri_tests = 0; res->num_contact = 0; res->UpboundTOC = 1; res->numCA = 0;
  res->toc = 0; res->collisionfree = false;
  if (!o1 || !o2 || !o1->gpu || !o2->gpu)
  {
    fprintf(stderr, "c2a_b200: C2A_QueryTimeOfContact on a model without EndModel()\n");
    res->numCA = -1;
    return 0;
  }
  double rec[48];
  record_from_motion(objmotion1, rec);
  record_from_motion(objmotion2, rec + 24);
  int32_t sa = seed_index(o1, res->last_triA), sb = seed_index(o2, res->last_triB);
  int32_t status = -1, cf = 0, nca = 0, nbv = 0, ntri = 0, last[2] = {-1, -1};
  double toc = 0, dist = 0, mint = 0, p1p2[6] = {0, 0, 0, 0, 0, 0};
  c2a_b200_results out;
  memset(&out, 0, sizeof(out));
  out.status = &status; out.collisionfree = &cf; out.num_ca = &nca; out.num_bv_tests = &nbv; out.num_tri_tests = &ntri;
  out.toc = &toc; out.distance = &dist; out.mint = &mint; out.p1p2 = p1p2; out.last_tri = last;
  const int rc = c2a_b200_solve_batch_motions(o1->gpu, o2->gpu, rec, &sa, &sb, 1, tolerance_d, tolerance_t, &out);
  if (rc != 0 || status != C2A_B200_QUERY_OK)
  {
    if (rc) fprintf(stderr, "c2a_b200: %s\n", c2a_b200_last_error());
    else fprintf(stderr, "c2a_b200: translation-only query on hierarchies deeper than its traversal stack\n");
    res->numCA = -1;
    return 0;
  }
  res->collisionfree = cf != 0;
  res->toc = toc; res->distance = dist; res->mint = mint; res->numCA = nca;
  res->num_bv_tests = nbv; res->num_tri_tests = ntri;
  for (int i = 0; i < 3; i++) { res->p1[i] = p1p2[i]; res->p2[i] = p1p2[3 + i]; }
  if (objmotion1->m_angVel < 1e-8 && objmotion2->m_angVel < 1e-8)
  {
    // translation-only branch (C2A.cpp:2032-2049): the traversal updates res->last_triA/B instead of the models'
    // last_tri (:1413-1414) and leaves both motions at the pose of the step bound (:1907-1908); numCA stays 0
    if (last[0] >= 0) res->last_triA = &o1->tris[last[0]];
    if (last[1] >= 0) res->last_triB = &o2->tris[last[1]];
    PQP_REAL Rt[3][3], Tt[3];
    objmotion1->integrate(mint, Rt, Tt);
    objmotion2->integrate(mint, Rt, Tt);
    return toc;
  }
  // the traversal's side effect on the models (C2A.cpp:1175-1176); the demo feeds it back as the next seeds
  if (last[0] >= 0) o1->last_tri = &o1->tris[last[0]];
  if (last[1] >= 0) o2->last_tri = &o2->tris[last[1]];
  if (!res->collisionfree) objmotion1->integrate(toc, res->R_toc, res->T_toc);  // C2A.cpp:2143
  return toc;
}

// C2A/src/C2A_PQP.cpp:970-1056
// C2A_Distance and C2A_Collide's C2A_DistanceResult overload: the same call but for the entry point
static int distance_like(bool gate, C2A_DistanceResult *res, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                         PQP_REAL T2[3], C2A_Model *o2, PQP_REAL rel_err, PQP_REAL abs_err, int qsize)
{
  if (!o1 || !o2 || !o1->gpu || !o2->gpu) return PQP_ERR_UNPROCESSED_MODEL;
  const auto t_begin = std::chrono::steady_clock::now();
  double pose[24];
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) { pose[3 * i + j] = R1[i][j]; pose[12 + 3 * i + j] = R2[i][j]; }
    pose[9 + i] = T1[i]; pose[21 + i] = T2[i];
  }
  int32_t sa = seed_index(o1, o1->last_tri), sb = seed_index(o2, o2->last_tri), pair[2] = {0, 0}, nbv = 0, ntri = 0;
  double dist = 0, p1p2[6];
  const int rc = gate ? c2a_b200_collide_distance_batch(o1->gpu, o2->gpu, pose, &sa, &sb, 1, rel_err, abs_err, &dist, p1p2, pair, &nbv, &ntri)
                      : c2a_b200_distance_queue_batch(o1->gpu, o2->gpu, pose, &sa, &sb, 1, rel_err, abs_err, qsize, &dist, p1p2, pair, &nbv, &ntri);
  if (rc) { fprintf(stderr, "c2a_b200: %s\n", c2a_b200_last_error()); return rc; }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) res->R[i][j] = (R1[0][i] * R2[0][j] + R1[1][i] * R2[1][j] + R1[2][i] * R2[2][j]);
  {
    const PQP_REAL Tt[3] = {T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
    for (int i = 0; i < 3; i++) res->T[i] = (R1[0][i] * Tt[0] + R1[1][i] * Tt[1] + R1[2][i] * Tt[2]);
  }
  res->distance = dist; res->rel_err = rel_err; res->abs_err = abs_err; res->qsize = qsize;
  res->num_bv_tests = nbv; res->num_tri_tests = ntri;
  for (int i = 0; i < 3; i++) { res->p1[i] = p1p2[i]; res->p2[i] = p1p2[3 + i]; }
  o1->last_tri = &o1->tris[pair[0]];
  o2->last_tri = &o2->tris[pair[1]];
  res->t1 = o1->tris[pair[0]].id; res->t2 = o2->tris[pair[1]].id;
  res->query_time_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  return PQP_OK;
}

int C2A_Distance(C2A_DistanceResult *res, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                 PQP_REAL T2[3], C2A_Model *o2, PQP_REAL rel_err, PQP_REAL abs_err, int qsize)
{
  return distance_like(false, res, R1, T1, o1, R2, T2, o2, rel_err, abs_err, qsize);
}

// C2A/src/C2A_PQP.cpp:1199-1280
int C2A_Collide(C2A_DistanceResult *res, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                PQP_REAL T2[3], C2A_Model *o2, PQP_REAL rel_err, PQP_REAL abs_err, int qsize)
{
  return distance_like(true, res, R1, T1, o1, R2, T2, o2, rel_err, abs_err, qsize);
}

void PQP_CollideResult::SizeTo(int n)
{
  if (n < num_pairs) return;
  CollisionPair *t = new CollisionPair[n];
  for (int i = 0; i < num_pairs; i++) t[i] = pairs[i];
  delete[] pairs;
  pairs = t;
  num_pairs_alloced = n;
}

void PQP_CollideResult::Add(int a, int b)
{
  if (num_pairs >= num_pairs_alloced) SizeTo(num_pairs_alloced * 2 + 8);
  pairs[num_pairs].id1 = a; pairs[num_pairs].id2 = b;
  num_pairs++;
}

// C2A/src/C2A_PQP.cpp:910-968.  The pair list is fetched with room for 256 pairs and once more with room for all of
// them when there are more.
int C2A_Collide(PQP_CollideResult *res, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3], PQP_REAL T2[3],
                C2A_Model *o2, int flag)
{
  if (!o1 || !o2 || !o1->gpu || !o2->gpu) return PQP_ERR_UNPROCESSED_MODEL;
  const auto t_begin = std::chrono::steady_clock::now();
  double pose[24];
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) { pose[3 * i + j] = R1[i][j]; pose[12 + 3 * i + j] = R2[i][j]; }
    pose[9 + i] = T1[i]; pose[21 + i] = T2[i];
  }
  int32_t cap = 256, n = 0, nbv = 0, ntri = 0;
  std::vector<int32_t> buf((size_t)2 * cap);
  int rc = c2a_b200_collide_batch(o1->gpu, o2->gpu, pose, 1, flag, cap, &n, buf.data(), &nbv, &ntri);
  if (rc == 0 && n > cap)
  {
    cap = n;
    buf.resize((size_t)2 * cap);
    rc = c2a_b200_collide_batch(o1->gpu, o2->gpu, pose, 1, flag, cap, &n, buf.data(), &nbv, &ntri);
  }
  if (rc) { fprintf(stderr, "c2a_b200: %s\n", c2a_b200_last_error()); return rc; }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) res->R[i][j] = (R1[0][i] * R2[0][j] + R1[1][i] * R2[1][j] + R1[2][i] * R2[2][j]);
  {
    const PQP_REAL Tt[3] = {T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
    for (int i = 0; i < 3; i++) res->T[i] = (R1[0][i] * Tt[0] + R1[1][i] * Tt[1] + R1[2][i] * Tt[2]);
  }
  res->num_bv_tests = nbv; res->num_tri_tests = ntri;
  res->num_pairs = 0;   // the reference keeps the allocation and resets the counter (:931)
  for (int k = 0; k < n; k++) res->Add(o1->tris[buf[2 * k]].id, o2->tris[buf[2 * k + 1]].id);
  res->query_time_secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  return PQP_OK;
}

// C2A/src/C2A.cpp:1937-1966: contact features at the motions' CURRENT poses
PQP_REAL C2A_QueryContact(CInterpMotion *objmotion1, CInterpMotion *objmotion2, C2A_TimeOfContactResult *res, C2A_Model *o1,
                          C2A_Model *o2, double threshold)
{
  res->num_contact = 0;
  res->UpboundTOC = 1;
  if (!o1 || !o2 || !o1->gpu || !o2->gpu) return 0;
  double poses[24];
  pose12(objmotion1->transform, poses);
  pose12(objmotion2->transform, poses + 12);
  const int cap = 256;
  std::vector<c2a_b200_contact> recs(cap);
  int32_t n = 0;
  int rc = c2a_b200_contacts_batch(o1->gpu, o2->gpu, poses, &threshold, 1, cap, &n, recs.data());
  if (rc == 0 && n > cap)
  {
    recs.resize(n);
    rc = c2a_b200_contacts_batch(o1->gpu, o2->gpu, poses, &threshold, 1, n, &n, recs.data());
  }
  if (rc != 0) { fprintf(stderr, "c2a_b200: %s\n", c2a_b200_last_error()); return 0; }
  // the reference push_front()s in visiting order
  for (int i = 0; i < n; i++)
  {
    const c2a_b200_contact &c = recs[i];
    ContactF f;
    f.FeatureType_A = c.type_a; f.FeatureType_B = c.type_b;
    for (int k = 0; k < 3; k++) { f.FeatureID_A[k] = c.fid_a[k]; f.FeatureID_B[k] = c.fid_b[k]; f.P_A[k] = c.pa[k]; f.P_B[k] = c.pb[k]; }
    f.TriangleID_A = c.tri_a; f.TriangleID_B = c.tri_b;
    f.Distance = c.dist;
    res->cont_l.push_front(f);
  }
  res->num_contact = n;
  return 0;
}

// C2A/src/C2A.cpp:1969-1985: the contact pass at explicit poses; unlike C2A_QueryContact it clears the list first
PQP_REAL C2A_QueryContactOnly(C2A_TimeOfContactResult *res, PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3],
                              PQP_REAL T2[3], C2A_Model *o2, double threshold)
{
  res->cont_l.clear();
  res->num_contact = 0;
  res->UpboundTOC = 1;
  if (!o1 || !o2 || !o1->gpu || !o2->gpu) return PQP_OK;
  double poses[24];
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) { poses[3 * i + j] = R1[i][j]; poses[12 + 3 * i + j] = R2[i][j]; }
    poses[9 + i] = T1[i]; poses[21 + i] = T2[i];
  }
  int cap = 256;
  std::vector<c2a_b200_contact> recs(cap);
  int32_t n = 0;
  int rc = c2a_b200_contacts_batch(o1->gpu, o2->gpu, poses, &threshold, 1, cap, &n, recs.data());
  if (rc == 0 && n > cap)
  {
    recs.resize(n);
    rc = c2a_b200_contacts_batch(o1->gpu, o2->gpu, poses, &threshold, 1, n, &n, recs.data());
  }
  if (rc != 0) { fprintf(stderr, "c2a_b200: %s\n", c2a_b200_last_error()); return PQP_OK; }
  for (int i = 0; i < n; i++)
  {
    const c2a_b200_contact &c = recs[i];
    ContactF f;
    f.FeatureType_A = c.type_a; f.FeatureType_B = c.type_b;
    for (int k = 0; k < 3; k++) { f.FeatureID_A[k] = c.fid_a[k]; f.FeatureID_B[k] = c.fid_b[k]; f.P_A[k] = c.pa[k]; f.P_B[k] = c.pb[k]; }
    f.TriangleID_A = c.tri_a; f.TriangleID_B = c.tri_b;
    f.Distance = c.dist;
    res->cont_l.push_front(f);
  }
  res->num_contact = n;
  return PQP_OK;
}

// C2A/src/C2A.cpp:1778-1931 (rotational branch)
int C2A_TimeOfContactStep(CInterpMotion *objmotion1, CInterpMotion *objmotion2, C2A_TimeOfContactResult *res,
                          PQP_REAL R1[3][3], PQP_REAL T1[3], C2A_Model *o1, PQP_REAL R2[3][3], PQP_REAL T2[3],
                          C2A_Model *o2, PQP_REAL tolerance_t, PQP_REAL tolerance_d)
{
  if (!o1 || !o2 || !o1->gpu || !o2->gpu) return PQP_ERR_UNPROCESSED_MODEL;
  double rec[48], step[C2A_B200_STEP_IN_DOUBLES];
  record_from_motion(objmotion1, rec);
  record_from_motion(objmotion2, rec + 24);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { step[3 * i + j] = R1[i][j]; step[12 + 3 * i + j] = R2[i][j]; }
  for (int i = 0; i < 3; i++) { step[9 + i] = T1[i]; step[21 + i] = T2[i]; }
  step[24] = (double)res->numCA; step[25] = res->mint; step[26] = res->UpboundTOC; step[27] = 0;
  int32_t sa = seed_index(o1, res->last_triA), sb = seed_index(o2, res->last_triB);
  int32_t status = -1, nbv = 0, ntri = 0, last[2] = {-1, -1};
  double dist = 0, mint = 0, p1p2[6] = {0, 0, 0, 0, 0, 0};
  c2a_b200_results out;
  memset(&out, 0, sizeof(out));
  out.status = &status; out.num_bv_tests = &nbv; out.num_tri_tests = &ntri; out.distance = &dist; out.mint = &mint; out.p1p2 = p1p2;
  out.last_tri = last;
  const int rc = c2a_b200_toc_step_batch(o1->gpu, o2->gpu, rec, step, &sa, &sb, 1, tolerance_t, tolerance_d, &out);
  if (rc != 0 || status != C2A_B200_QUERY_OK) return rc ? rc : PQP_ERR_UNPROCESSED_MODEL;
  // res->R, res->T: the relative transform the step computes (C2A.cpp:1793-1796)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) res->R[i][j] = (R1[0][i] * R2[0][j] + R1[1][i] * R2[1][j] + R1[2][i] * R2[2][j]);
  {
    const PQP_REAL Tt[3] = {T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
    for (int i = 0; i < 3; i++) res->T[i] = (R1[0][i] * Tt[0] + R1[1][i] * Tt[1] + R1[2][i] * Tt[2]);
  }
  res->distance = dist; res->mint = mint;
  res->num_bv_tests += nbv; res->num_tri_tests += ntri;
  if (last[0] >= 0) o1->last_tri = &o1->tris[last[0]];
  if (last[1] >= 0) o2->last_tri = &o2->tris[last[1]];
  if (ntri > 0 && (p1p2[0] != 0 || p1p2[1] != 0 || p1p2[2] != 0 || p1p2[3] != 0 || p1p2[4] != 0 || p1p2[5] != 0))
    for (int i = 0; i < 3; i++) { res->p1[i] = p1p2[i]; res->p2[i] = p1p2[3 + i]; }
  return PQP_OK;
}

// C2A/src/C2A.cpp:2315-2444 (without the contact pass)
C2A_Result C2A_Solve(Transform *trans00, Transform *trans01, C2A_Model *obj1_tested, Transform *trans10,
                     Transform *trans11, C2A_Model *obj2_tested, Transform &trans0, Transform &trans1,
                     PQP_REAL &time_of_contact, int &number_of_iteration, int &number_of_contact, PQP_REAL th_ca,
                     C2A_TimeOfContactResult &dres)
{
  (void)th_ca;  // ignored by the reference as well
  number_of_iteration = 0;
  PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3], R1e[3][3], T1e[3], R2e[3][3], T2e[3];
  trans00->Rotation().Get_Value(R1); trans00->Translation().Get_Value(T1);
  trans10->Rotation().Get_Value(R2); trans10->Translation().Get_Value(T2);
  trans01->Rotation().Get_Value(R1e); trans01->Translation().Get_Value(T1e);
  trans11->Rotation().Get_Value(R2e); trans11->Translation().Get_Value(T2e);
  CInterpMotion_Linear motion1(R1, T1, R1e, T1e);
  CInterpMotion_Linear motion2(R2, T2, R2e, T2e);
  const PQP_REAL t_delta = 0.0001, d_delta = 0.0001;  // hard-coded in the reference, C2A.cpp:2384-2385
  motion1.m_toc_delta = d_delta;
  motion2.m_toc_delta = d_delta;
  C2A_QueryTimeOfContact(&motion1, &motion2, &dres, obj1_tested, obj2_tested, d_delta, t_delta, 0);
  dres.cont_l.clear();
  if (dres.numCA < 0)
  {
    time_of_contact = 0; number_of_contact = 0; number_of_iteration = -1;
    return CollisionNotFound;
  }
  if (!dres.collisionfree)
  {
    PQP_REAL qua[7];
    motion1.integrate(dres.toc, qua);
    trans0.Set_Rotation(Quaternion(qua[1], qua[2], qua[3], qua[0]));
    trans0.Set_Translation(Coord3D(qua[4], qua[5], qua[6]));
    motion2.integrate(dres.toc, qua);
    trans1.Set_Rotation(Quaternion(qua[1], qua[2], qua[3], qua[0]));
    trans1.Set_Translation(Coord3D(qua[4], qua[5], qua[6]));
    const double threshold = 2 * dres.distance + 0.001;  // C2A.cpp:2433
    C2A_QueryContact(&motion1, &motion2, &dres, obj1_tested, obj2_tested, threshold);
  }
  time_of_contact = dres.toc;
  number_of_contact = dres.num_contact;
  number_of_iteration = dres.numCA;
  return TOCFound;
}

static int solve_batch_common(const int *devices, int n_devices, int n, const Transform *trans00, const Transform *trans01,
                              C2A_Model *obj1_tested, const Transform *trans10, const Transform *trans11, C2A_Model *obj2_tested,
                              const int *seed_tri_a, const int *seed_tri_b, bool *collisionfree, PQP_REAL *time_of_contact,
                              PQP_REAL *distance, int *number_of_iteration, Transform *trans0, Transform *trans1)
{
  if (n < 0 || !obj1_tested || !obj2_tested || !obj1_tested->gpu || !obj2_tested->gpu) return C2A_B200_ERR_ARG;
  if (n == 0) return PQP_OK;
  std::vector<double> poses((size_t)48 * n), pose_toc(trans0 || trans1 ? (size_t)24 * n : 0);
  for (int i = 0; i < n; i++)
  {
    pose12(trans00[i], &poses[(size_t)48 * i]); pose12(trans01[i], &poses[(size_t)48 * i + 12]);
    pose12(trans10[i], &poses[(size_t)48 * i + 24]); pose12(trans11[i], &poses[(size_t)48 * i + 36]);
  }
  std::vector<int32_t> cf(n), nca(n), status(n);
  c2a_b200_results out;
  memset(&out, 0, sizeof(out));
  out.status = status.data(); out.collisionfree = cf.data(); out.num_ca = nca.data();
  out.toc = time_of_contact; out.distance = distance;
  if (!pose_toc.empty()) out.pose_toc = pose_toc.data();
  int rc;
  if (!devices)
    rc = c2a_b200_solve_batch(obj1_tested->gpu, obj2_tested->gpu, poses.data(), seed_tri_a, seed_tri_b, n, 0.0001, 0.0001, &out);
  else
  {
    std::vector<const c2a_b200_model *> a(n_devices), b(n_devices);
    for (int d = 0; d < n_devices; d++)
    {
      a[d] = obj1_tested->OnDevice(devices[d]); b[d] = obj2_tested->OnDevice(devices[d]);
      if (!a[d] || !b[d]) return C2A_B200_ERR_DEVICE;  // no replica there: C2A_Model::ReplicateTo first
    }
    rc = c2a_b200_solve_batch_multi(a.data(), b.data(), n_devices, poses.data(), seed_tri_a, seed_tri_b, n, 0.0001, 0.0001, &out);
  }
  if (rc) return rc;
  for (int i = 0; i < n; i++)
  {
    const bool ok = status[i] == C2A_B200_QUERY_OK;
    if (collisionfree) collisionfree[i] = ok && cf[i] != 0;
    if (number_of_iteration) number_of_iteration[i] = ok ? nca[i] : -1;
    if (ok && !cf[i])
    {
      if (trans0) { trans0[i].Rotation().Set_Value(&pose_toc[(size_t)24 * i]); trans0[i].Translation().Set_Value(&pose_toc[(size_t)24 * i + 9]); }
      if (trans1) { trans1[i].Rotation().Set_Value(&pose_toc[(size_t)24 * i + 12]); trans1[i].Translation().Set_Value(&pose_toc[(size_t)24 * i + 21]); }
    }
  }
  return PQP_OK;
}

int C2A_SolveBatch(int n, const Transform *trans00, const Transform *trans01, C2A_Model *obj1_tested,
                   const Transform *trans10, const Transform *trans11, C2A_Model *obj2_tested,
                   const int *seed_tri_a, const int *seed_tri_b, bool *collisionfree, PQP_REAL *time_of_contact,
                   PQP_REAL *distance, int *number_of_iteration, Transform *trans0, Transform *trans1)
{
  return solve_batch_common(0, 0, n, trans00, trans01, obj1_tested, trans10, trans11, obj2_tested, seed_tri_a, seed_tri_b, collisionfree,
                            time_of_contact, distance, number_of_iteration, trans0, trans1);
}

int C2A_SolveBatchMulti(const int *devices, int n_devices, int n, const Transform *trans00, const Transform *trans01,
                        C2A_Model *obj1_tested, const Transform *trans10, const Transform *trans11, C2A_Model *obj2_tested,
                        const int *seed_tri_a, const int *seed_tri_b, bool *collisionfree, PQP_REAL *time_of_contact,
                        PQP_REAL *distance, int *number_of_iteration, Transform *trans0, Transform *trans1)
{
  if (!devices || n_devices <= 0) return C2A_B200_ERR_ARG;
  return solve_batch_common(devices, n_devices, n, trans00, trans01, obj1_tested, trans10, trans11, obj2_tested, seed_tri_a, seed_tri_b,
                            collisionfree, time_of_contact, distance, number_of_iteration, trans0, trans1);
}
