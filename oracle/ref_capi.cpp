// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
//
// C-ABI wrapper around the REFERENCE's own object code: /root/reference/C2A/src/*.cpp
// compiled verbatim (see oracle/Makefile) against oracle/pqp_shim.  Built into
// oracle/_ref/libc2a_ref.so.  Used by tests/ to pin the oracle port and the CUDA
// path, by tests/golden/make_golden.py to generate fixtures, and by bench.py as
// the CPU baseline ("kind": "reference").  No reference source is copied here;
// this file only calls the reference's public entry points.
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <thread>
#include <vector>

#include "PQP.h"
#include "MatVec.h"
#include "C2A/C2A.h"
#include "C2A/InterpMotion.h"
#include "C2A/LinearMath.h"
#include "C2A/C2A_RectDist.h"

#include "c2a_oracle.h"  // orc_result layout only

extern bool b_TanslationCCD;  // /root/reference/C2A/src/C2A.cpp:33

namespace {
struct StdoutSilencer
{
  int saved;
  StdoutSilencer()
  {
    fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~StdoutSilencer()
  {
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
  }
};

inline void pose_to_RT(const double *p, PQP_REAL R[3][3], PQP_REAL T[3])
{
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) R[i][j] = p[3 * i + j];
    T[i] = p[9 + i];
  }
}

inline void transform_to_pose(Transform &t, double *p)
{
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) p[3 * i + j] = t.Rotation()[i][j];
    p[9 + i] = t.Translation()[i];
  }
}

inline void fill_common(orc_result *o, C2A_TimeOfContactResult &dres)
{
  o->collisionfree = dres.collisionfree ? 1 : 0;
  o->numCA = dres.numCA;
  o->num_bv_tests = dres.num_bv_tests;
  o->num_tri_tests = dres.num_tri_tests;
  o->toc = dres.toc;
  o->distance = dres.distance;
  o->mint = dres.mint;
  for (int k = 0; k < 3; k++) { o->p1[k] = dres.p1[k]; o->p2[k] = dres.p2[k]; }
  o->last_tri_a = o->last_tri_b = -1;
}

// o->last_tri after a query (only meaningful when queries run one at a time: it is model state)
inline void fill_last_tri(orc_result *o, C2A_Model *A, C2A_Model *B)
{
  o->last_tri_a = A->last_tri ? (int)((C2A_Tri *)A->last_tri - (C2A_Tri *)A->tris) : -1;
  o->last_tri_b = B->last_tri ? (int)((C2A_Tri *)B->last_tri - (C2A_Tri *)B->tris) : -1;
}

// The body of C2A_Solve (C2A/src/C2A.cpp:2315-2444) without its printf and contact
// pass, with the tolerances exposed: calls the reference's CInterpMotion_Linear
// and C2A_QueryTimeOfContact.
void query_toc(C2A_Model *A, C2A_Model *B, const double *poses, int seedA, int seedB, double tol_d,
               double tol_t, orc_result *o, bool allow_translation)
{
  PQP_REAL R1[3][3], T1[3], R1e[3][3], T1e[3], R2[3][3], T2[3], R2e[3][3], T2e[3];
  pose_to_RT(poses + 0, R1, T1); pose_to_RT(poses + 12, R1e, T1e);
  pose_to_RT(poses + 24, R2, T2); pose_to_RT(poses + 36, R2e, T2e);
  CInterpMotion_Linear motion1(R1, T1, R1e, T1e);
  CInterpMotion_Linear motion2(R2, T2, R2e, T2e);
  motion1.m_toc_delta = tol_d;
  motion2.m_toc_delta = tol_d;
  memset(o, 0, sizeof(*o));
  bool translation = (motion1.m_angVel < 1e-8 && motion2.m_angVel < 1e-8);
  if (translation && !allow_translation) { o->collisionfree = -2; return; }  // deferred to the serial phase
  b_TanslationCCD = translation;
  C2A_TimeOfContactResult dres;
  dres.last_triA = A->GetTriangle(seedA);
  dres.last_triB = B->GetTriangle(seedB);
  if (allow_translation) { A->last_tri = 0; B->last_tri = 0; }  // serial mode: observe the traversal's writes
  C2A_QueryTimeOfContact(&motion1, &motion2, &dres, A, B, tol_d, tol_t, 0);
  fill_common(o, dres);
  if (allow_translation) fill_last_tri(o, A, B);
  if (translation)
  {
    // the translation-only traversal updates res->last_triA/B instead of o->last_tri (C2A.cpp:1413-1414)
    o->last_tri_a = (int)((C2A_Tri *)dres.last_triA - (C2A_Tri *)A->tris);
    o->last_tri_b = (int)((C2A_Tri *)dres.last_triB - (C2A_Tri *)B->tris);
  }
  if (!dres.collisionfree)
  {
    PQP_REAL qua[7];
    motion1.integrate(dres.toc, qua);
    motion2.integrate(dres.toc, qua);
    transform_to_pose(motion1.transform, &o->pose_toc[0]);
    transform_to_pose(motion2.transform, &o->pose_toc[12]);
  }
}

void full_solve(C2A_Model *A, C2A_Model *B, const double *poses, int seedA, int seedB, orc_result *o,
                int *num_contact)
{
  Transform t00, t01, t10, t11, out0, out1;
  Real v[12];
  Transform *tt[4] = {&t00, &t01, &t10, &t11};
  for (int k = 0; k < 4; k++)
  {
    const double *p = poses + 12 * k;
    for (int i = 0; i < 3; i++) { v[4 * i + 0] = p[3 * i]; v[4 * i + 1] = p[3 * i + 1]; v[4 * i + 2] = p[3 * i + 2]; v[4 * i + 3] = p[9 + i]; }
    tt[k]->Set_Value(v);
  }
  out0.Identity(); out1.Identity();
  C2A_TimeOfContactResult dres;
  dres.last_triA = A->GetTriangle(seedA);
  dres.last_triB = B->GetTriangle(seedB);
  PQP_REAL toc = 0;
  int nIter = 0, nContact = 0;
  C2A_Solve(&t00, &t01, A, &t10, &t11, B, out0, out1, toc, nIter, nContact, 0.0, dres);
  memset(o, 0, sizeof(*o));
  fill_common(o, dres);
  if (!dres.collisionfree)
  {
    transform_to_pose(out0, &o->pose_toc[0]);
    transform_to_pose(out1, &o->pose_toc[12]);
  }
  if (num_contact) *num_contact = nContact;
}
}  // namespace

extern "C" {

// tris9: [n_tris][9] = p1,p2,p3.  Mirrors the demo's loader loop
// (/root/reference/CCDDemo/mainTorusknot.cpp:448-485): BeginModel, AddTri(p1,p2,p3,i,i1,i2,i3), EndModel.
void *ref_model_build(const double *tris9, const int32_t *vidx3, int32_t n_tris)
{
  StdoutSilencer quiet;
  C2A_Model *m = new C2A_Model;
  m->BeginModel();
  for (int i = 0; i < n_tris; i++)
  {
    const double *t = tris9 + 9 * i;
    int i1 = vidx3 ? vidx3[3 * i] : 3 * i, i2 = vidx3 ? vidx3[3 * i + 1] : 3 * i + 1, i3 = vidx3 ? vidx3[3 * i + 2] : 3 * i + 2;
    m->AddTri(t, t + 3, t + 6, i, i1, i2, i3);
  }
  m->EndModel();
  return m;
}

void ref_model_free(void *h) { delete (C2A_Model *)h; }

void ref_model_counts(void *h, int32_t *n_nodes, int32_t *n_tris)
{
  C2A_Model *m = (C2A_Model *)h;
  *n_nodes = m->num_bvs;
  *n_tris = m->num_tris;
}

// Flatten the hot fields of the built model.
void ref_model_export(void *h, double *R, double *Tr, double *l, double *r, double *R_loc, double *ang,
                      int32_t *first_child, double *tris9, int32_t *tri_ids)
{
  C2A_Model *m = (C2A_Model *)h;
  for (int n = 0; n < m->num_bvs; n++)
  {
    C2A_BV *b = (C2A_BV *)m->child(n);
    for (int i = 0; i < 3; i++)
    {
      for (int j = 0; j < 3; j++) { R[9 * n + 3 * i + j] = b->R[i][j]; R_loc[9 * n + 3 * i + j] = b->R_loc[i][j]; }
      Tr[3 * n + i] = b->Tr[i];
    }
    l[2 * n] = b->l[0]; l[2 * n + 1] = b->l[1];
    r[n] = b->r;
    ang[n] = b->angularRadius;
    first_child[n] = b->first_child;
  }
  for (int t = 0; t < m->num_tris; t++)
  {
    C2A_Tri *tr = m->GetTriangle(t);
    for (int k = 0; k < 3; k++) { tris9[9 * t + k] = tr->p1[k]; tris9[9 * t + 3 + k] = tr->p2[k]; tris9[9 * t + 6 + k] = tr->p3[k]; }
    if (tri_ids) tri_ids[t] = tr->id;
  }
}

// mode 0: C2A_Solve's TOC part (motions + C2A_QueryTimeOfContact + pose at TOC), tolerances exposed.
// mode 1: the full, unmodified C2A_Solve (tolerances hard-coded 1e-4; includes printf + contact pass);
//         num_contact[n] optional.
void ref_solve_batch(void *hA, void *hB, const double *poses, int64_t n, const int32_t *seedA,
                     const int32_t *seedB, int32_t mode, double tol_d, double tol_t, orc_result *out,
                     int32_t *num_contact, int32_t n_threads)
{
  C2A_Model *A = (C2A_Model *)hA, *B = (C2A_Model *)hB;
  StdoutSilencer quiet;
  if (n_threads < 1) n_threads = 1;
  if (mode == 1 || n_threads == 1)
  {
    // C2A_Solve writes the global b_TanslationCCD per call (C2A.cpp:2391-2395) and printf-locks stdout:
    // run it on one thread so the reference stays unmodified and race-free.
    for (int64_t i = 0; i < n; i++)
    {
      int sa = seedA ? seedA[i] : 0, sb = seedB ? seedB[i] : 0;
      if (mode == 1) full_solve(A, B, poses + 48 * i, sa, sb, &out[i], num_contact ? &num_contact[i] : 0);
      else query_toc(A, B, poses + 48 * i, sa, sb, tol_d, tol_t, &out[i], true);
    }
    return;
  }
  // threaded phase: rotational queries only (b_TanslationCCD stays false, written with the same value)
  b_TanslationCCD = false;
  // dynamic queue of 16-query chunks: the cost per query has a long tail, a static split would time the unluckiest thread
  std::vector<std::thread> th;
  std::atomic<int64_t> next(0);
  for (int t = 0; t < n_threads; t++)
    th.emplace_back([=, &next]() {
      while (true)
      {
        const int64_t lo = next.fetch_add(16);
        if (lo >= n) break;
        const int64_t hi = lo + 16 < n ? lo + 16 : n;
        for (int64_t i = lo; i < hi; i++)
          query_toc(A, B, poses + 48 * i, seedA ? seedA[i] : 0, seedB ? seedB[i] : 0, tol_d, tol_t, &out[i], false);
      }
    });
  for (auto &t : th) t.join();
  // serial phase: translation-only queries
  for (int64_t i = 0; i < n; i++)
    if (out[i].collisionfree == -2)
      query_toc(A, B, poses + 48 * i, seedA ? seedA[i] : 0, seedB ? seedB[i] : 0, tol_d, tol_t, &out[i], true);
  b_TanslationCCD = false;
}

// The full, unmodified C2A_Solve for one query, with its contact list exported.  The reference
// push_front()s contacts, so list order = reverse visiting order; records are written in LIST order.
// FeatureID entries the reference leaves uninitialised are copied as they are (garbage): compare only
// the first 1 / 2 / 3 entries for a vertex / edge / face feature.
int64_t ref_solve_contacts(void *hA, void *hB, const double *poses48, int32_t seedA, int32_t seedB, orc_result *res,
                           int64_t max_out, orc_contact *out)
{
  C2A_Model *A = (C2A_Model *)hA, *B = (C2A_Model *)hB;
  StdoutSilencer quiet;
  Transform t00, t01, t10, t11, out0, out1;
  Real v[12];
  Transform *tt[4] = {&t00, &t01, &t10, &t11};
  for (int k = 0; k < 4; k++)
  {
    const double *p = poses48 + 12 * k;
    for (int i = 0; i < 3; i++) { v[4 * i + 0] = p[3 * i]; v[4 * i + 1] = p[3 * i + 1]; v[4 * i + 2] = p[3 * i + 2]; v[4 * i + 3] = p[9 + i]; }
    tt[k]->Set_Value(v);
  }
  out0.Identity(); out1.Identity();
  C2A_TimeOfContactResult dres;
  dres.last_triA = A->GetTriangle(seedA);
  dres.last_triB = B->GetTriangle(seedB);
  PQP_REAL toc = 0;
  int nIter = 0, nContact = 0;
  C2A_Solve(&t00, &t01, A, &t10, &t11, B, out0, out1, toc, nIter, nContact, 0.0, dres);
  memset(res, 0, sizeof(*res));
  fill_common(res, dres);
  if (!dres.collisionfree)
  {
    transform_to_pose(out0, &res->pose_toc[0]);
    transform_to_pose(out1, &res->pose_toc[12]);
  }
  int64_t k = 0;
  for (ContactFListIterator it = dres.cont_l.begin(); it != dres.cont_l.end(); ++it, ++k)
  {
    if (k >= max_out) continue;
    orc_contact &c = out[k];
    c.type_a = it->FeatureType_A; c.type_b = it->FeatureType_B;
    for (int i = 0; i < 3; i++) { c.fid_a[i] = it->FeatureID_A[i]; c.fid_b[i] = it->FeatureID_B[i]; c.pa[i] = it->P_A[i]; c.pb[i] = it->P_B[i]; }
    c.tri_a = it->TriangleID_A; c.tri_b = it->TriangleID_B;
    c.dist = it->Distance;
  }
  return nContact;
}

// The reference's C2A_Distance (C2A/src/C2A_PQP.cpp:970-1056) for one query; seeds go in through o->last_tri as
// the reference expects, the closest pair comes back as builder-order indices (o->last_tri after the call).
void ref_distance(void *hA, void *hB, const double *pose24, int32_t seedA, int32_t seedB, double rel_err, double abs_err,
                  int32_t qsize, orc_distance_result *out)
{
  C2A_Model *A = (C2A_Model *)hA, *B = (C2A_Model *)hB;
  PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3];
  pose_to_RT(pose24, R1, T1); pose_to_RT(pose24 + 12, R2, T2);
  A->last_tri = A->GetTriangle(seedA);
  B->last_tri = B->GetTriangle(seedB);
  C2A_DistanceResult res;
  C2A_Distance(&res, R1, T1, A, R2, T2, B, rel_err, abs_err, qsize);
  out->distance = res.distance;
  for (int k = 0; k < 3; k++) { out->p1[k] = res.p1[k]; out->p2[k] = res.p2[k]; }
  out->tri_a = (int)((C2A_Tri *)A->last_tri - (C2A_Tri *)A->tris);
  out->tri_b = (int)((C2A_Tri *)B->last_tri - (C2A_Tri *)B->tris);
  out->num_bv_tests = res.num_bv_tests; out->num_tri_tests = res.num_tri_tests;
}

// OBB fields of the built model (BV::d, BV::To after make_parent_relative), read by C2A_Collide only.
void ref_model_export_obb(void *h, double *d, double *To)
{
  C2A_Model *m = (C2A_Model *)h;
  for (int n = 0; n < m->num_bvs; n++)
  {
    BV *b = m->child(n);
    for (int i = 0; i < 3; i++) { d[3 * n + i] = b->d[i]; To[3 * n + i] = b->To[i]; }
  }
}

// The reference's C2A_Collide(PQP_CollideResult*, ...) (C2A/src/C2A_PQP.cpp:910-968) for one query.  pairs: the first
// min(num_pairs, max_pairs) (Id1, Id2) = AddTri ids, in the order the traversal reported them.  Returns num_pairs.
int32_t ref_collide(void *hA, void *hB, const double *pose24, int32_t flag, int32_t max_pairs, int32_t *pairs,
                    int32_t *num_bv_tests, int32_t *num_tri_tests)
{
  C2A_Model *A = (C2A_Model *)hA, *B = (C2A_Model *)hB;
  PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3];
  pose_to_RT(pose24, R1, T1); pose_to_RT(pose24 + 12, R2, T2);
  PQP_CollideResult res;
  C2A_Collide(&res, R1, T1, A, R2, T2, B, flag);
  const int n = res.NumPairs();
  for (int k = 0; k < n && k < max_pairs; k++) { pairs[2 * k] = res.Id1(k); pairs[2 * k + 1] = res.Id2(k); }
  if (num_bv_tests) *num_bv_tests = res.NumBVTests();
  if (num_tri_tests) *num_tri_tests = res.NumTriTests();
  return n;
}

// The reference's C2A_Collide(C2A_DistanceResult*, ...) (C2A/src/C2A_PQP.cpp:1199-1280): the distance walk restricted to
// node pairs whose boxes overlap.  Seeds and closest pair as in ref_distance.
void ref_collide_distance(void *hA, void *hB, const double *pose24, int32_t seedA, int32_t seedB, double rel_err,
                          double abs_err, orc_distance_result *out)
{
  C2A_Model *A = (C2A_Model *)hA, *B = (C2A_Model *)hB;
  PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3];
  pose_to_RT(pose24, R1, T1); pose_to_RT(pose24 + 12, R2, T2);
  A->last_tri = A->GetTriangle(seedA);
  B->last_tri = B->GetTriangle(seedB);
  C2A_DistanceResult res;
  C2A_Collide(&res, R1, T1, A, R2, T2, B, rel_err, abs_err, 2);
  out->distance = res.distance;
  for (int k = 0; k < 3; k++) { out->p1[k] = res.p1[k]; out->p2[k] = res.p2[k]; }
  out->tri_a = (int)((C2A_Tri *)A->last_tri - (C2A_Tri *)A->tris);
  out->tri_b = (int)((C2A_Tri *)B->last_tri - (C2A_Tri *)B->tris);
  out->num_bv_tests = res.num_bv_tests; out->num_tri_tests = res.num_tri_tests;
}

// ---- unit-level entry points for pinning the port / the device functions ----
double ref_rect_dist(const double Rab[9], const double Tab[3], const double a[2], const double b[2], double P[3],
                     double Q[3], double S[3])
{
  PQP_REAL R[3][3], T[3], aa[2] = {a[0], a[1]}, bb[2] = {b[0], b[1]};
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i][j] = Rab[3 * i + j]; T[i] = Tab[i]; }
  bool valid;
  return C2ARectDist(R, T, aa, bb, P, Q, valid, S);
}

// the reference's in-tree triangle distance, C2A/src/C2A.cpp:408-424
double ref_tri_distance_intree(const double R9[9], const double T3[3], const double t1[9], const double t2[9],
                               double p[3], double q[3], int32_t *collided)
{
  PQP_REAL R[3][3], T[3];
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i][j] = R9[3 * i + j]; T[i] = T3[i]; }
  Tri a, b;
  memcpy(a.p1, t1, 24); memcpy(a.p2, t1 + 3, 24); memcpy(a.p3, t1 + 6, 24);
  memcpy(b.p1, t2, 24); memcpy(b.p2, t2 + 3, 24); memcpy(b.p3, t2 + 6, 24);
  C2A_ContactFeature f1, f2;
  bool bc = false;
  double d = C2A_TriDistance(R, T, &a, &b, p, q, f1, f2, bc);
  if (collided) *collided = bc ? 1 : 0;
  return d;
}

// CInterpMotion_Linear: constants, pose at t, and the two motion bounds.
// out = cv(3) axis(3) angVel(1) R(9) T(3) bound_bv(1) bound_leaf(1)  (19+2 doubles)
void ref_motion_probe(const double poses24[24], double t, double ang_radius, const double dir[3], double out[21])
{
  PQP_REAL R0[3][3], T0[3], R1[3][3], T1[3];
  pose_to_RT(poses24, R0, T0); pose_to_RT(poses24 + 12, R1, T1);
  CInterpMotion_Linear m(R0, T0, R1, T1);
  m.cv.Get_Value(out + 0);
  m.m_axis.Get_Value(out + 3);
  out[6] = m.m_angVel;
  PQP_REAL R[3][3], T[3];
  ((CInterpMotion *)&m)->integrate(t, R, T);  // base overload, as C2A.cpp:2112 calls it
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) out[7 + 3 * i + j] = R[i][j]; out[16 + i] = T[i]; }
  C2A_BV bv;
  bv.angularRadius = ang_radius;
  PQP_REAL n1[3] = {dir[0], dir[1], dir[2]}, n2[3] = {dir[0], dir[1], dir[2]}, Tdummy[3] = {0, 0, 0};
  out[19] = m.computeTOC_MotionBound(Tdummy, 1.0, &bv, n1);
  out[20] = m.computeTOC(1.0, ang_radius, n2);
}

}  // extern "C"
