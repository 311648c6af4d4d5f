// Drop-in check: this program is written against the REFERENCE's API (the way CCDDemo/mainTorusknot.cpp
// uses it: load a .tri mesh, BeginModel/AddTri/EndModel, per frame C2A_Solve with dres.last_triA/B seeded
// from the model) and compiles unchanged against include/C2A of this repo + libc2a_b200.so.
// It also drives C2A_QueryTimeOfContact, C2A_TimeOfContactStep and C2A_SolveBatch directly.
// Usage: dropin_demo mesh.txt poses.txt nframes   (plain-text inputs written by tests/test_gpu_dropin.py)
// Output: one line per frame: collisionfree toc distance numCA nbv ntri (hex floats), then cross-checks.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "PQP.h"
#include "C2A/C2A.h"
#include "C2A/LinearMath.h"
#include "C2A/InterpMotion.h"

static void set_transform(Transform &t, const double *p)
{
  for (int i = 0; i < 3; i++)
  {
    for (int j = 0; j < 3; j++) t.Rotation()[i][j] = p[3 * i + j];
    t.Translation()[i] = p[9 + i];
  }
}

int main(int argc, char **argv)
{
  if (argc < 4) return 2;
  FILE *fm = fopen(argv[1], "r");
  int nv = 0, nt = 0;
  if (!fm || fscanf(fm, "%d %d", &nv, &nt) != 2) return 3;
  std::vector<double> p(3 * nv);
  for (int i = 0; i < 3 * nv; i++) if (fscanf(fm, "%lf", &p[i]) != 1) return 3;
  C2A_Model *object1_tested = new C2A_Model(), *object2_tested = new C2A_Model();
  object1_tested->BeginModel();
  object2_tested->BeginModel();
  for (int i = 0; i < nt; i++)
  {
    int i1, i2, i3;
    if (fscanf(fm, "%d %d %d", &i1, &i2, &i3) != 3) return 3;
    object1_tested->AddTri(&p[3 * i1], &p[3 * i2], &p[3 * i3], i, i1, i2, i3);
    object2_tested->AddTri(&p[3 * i1], &p[3 * i2], &p[3 * i3], i, i1, i2, i3);
  }
  fclose(fm);
  if (object1_tested->EndModel() != PQP_OK || object2_tested->EndModel() != PQP_OK) return 4;

  const int nframes = atoi(argv[3]);
  FILE *fp = fopen(argv[2], "r");
  std::vector<double> poses(48 * nframes);
  for (int i = 0; i < 48 * nframes; i++) if (fscanf(fp, "%lf", &poses[i]) != 1) return 5;
  fclose(fp);

  C2A_TimeOfContactResult dres;
  std::vector<Transform> t00(nframes), t01(nframes), t10(nframes), t11(nframes);
  std::vector<double> toc_single(nframes), dist_single(nframes);
  std::vector<int> free_single(nframes), it_single(nframes), seed_a(nframes), seed_b(nframes);
  for (int f = 0; f < nframes; f++)
  {
    Transform trans0, trans1;
    set_transform(t00[f], &poses[48 * f]); set_transform(t01[f], &poses[48 * f + 12]);
    set_transform(t10[f], &poses[48 * f + 24]); set_transform(t11[f], &poses[48 * f + 36]);
    PQP_REAL toc; int nItr, NTr;
    dres.last_triA = object1_tested->last_tri;  // the demo carries the models' last_tri into every call
    dres.last_triB = object2_tested->last_tri;
    seed_a[f] = (int)((C2A_Tri *)dres.last_triA - object1_tested->tris);
    seed_b[f] = (int)((C2A_Tri *)dres.last_triB - object2_tested->tris);
    C2A_Result r = C2A_Solve(&t00[f], &t01[f], object1_tested, &t10[f], &t11[f], object2_tested, trans0, trans1, toc, nItr, NTr,
                             0.0, dres);
    if (r != TOCFound) return 6;
    printf("F %d %a %a %d %d %d %d %d", dres.collisionfree ? 1 : 0, toc, dres.Distance(), nItr, dres.NumBVTests(), dres.NumTriTests(), NTr,
           (int)dres.cont_l.size());
    if (NTr > 0) printf(" %d %d %a", dres.cont_l.front().TriangleID_A, dres.cont_l.front().TriangleID_B, dres.cont_l.front().Distance);
    else printf(" -1 -1 0x0p+0");
    if (!dres.collisionfree)
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) printf(" %a", trans0.Rotation()[i][j]);
    printf("\n");
    toc_single[f] = toc; dist_single[f] = dres.Distance(); free_single[f] = dres.collisionfree; it_single[f] = nItr;
  }

  // C2A_SolveBatch must agree with the per-frame calls
  std::vector<char> cfree(nframes);
  std::vector<double> tocs(nframes), dists(nframes);
  std::vector<int> its(nframes);
  bool *cf = new bool[nframes];
  if (C2A_SolveBatch(nframes, t00.data(), t01.data(), object1_tested, t10.data(), t11.data(), object2_tested, seed_a.data(), seed_b.data(), cf, tocs.data(),
                     dists.data(), its.data(), 0, 0) != PQP_OK) return 7;
  int batch_bad = 0;
  for (int f = 0; f < nframes; f++)
    if (cf[f] != (free_single[f] != 0) || tocs[f] != toc_single[f] || dists[f] != dist_single[f] || its[f] != it_single[f]) batch_bad++;
  printf("BATCH_MISMATCH %d\n", batch_bad);

  // the multi-device entry over the device(s) the models live on must agree as well (a box with more GPUs gets replicas)
  {
    int devs[2] = {object1_tested->device, object1_tested->device + 1};
    int nd = 1;
    if (object1_tested->ReplicateTo(devs[1]) == PQP_OK && object2_tested->ReplicateTo(devs[1]) == PQP_OK) nd = 2;
    std::vector<double> tocs2(nframes), dists2(nframes);
    std::vector<int> its2(nframes);
    bool *cf2 = new bool[nframes];
    if (C2A_SolveBatchMulti(devs, nd, nframes, t00.data(), t01.data(), object1_tested, t10.data(), t11.data(), object2_tested, seed_a.data(),
                            seed_b.data(), cf2, tocs2.data(), dists2.data(), its2.data(), 0, 0) != PQP_OK) return 11;
    int multi_bad = 0;
    for (int f = 0; f < nframes; f++)
      if (cf2[f] != cf[f] || tocs2[f] != tocs[f] || dists2[f] != dists[f] || its2[f] != its[f]) multi_bad++;
    printf("MULTI_MISMATCH %d devices %d\n", multi_bad, nd);
    delete[] cf2;
  }

  // C2A_QueryContactOnly at the first frames' start poses against C2A_QueryContact on motions standing at those poses
  {
    int only_bad = 0;
    for (int f = 0; f < nframes && f < 12; f++)
    {
      PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3];
      t00[f].Rotation().Get_Value(R1); t00[f].Translation().Get_Value(T1);
      t10[f].Rotation().Get_Value(R2); t10[f].Translation().Get_Value(T2);
      CInterpMotion_Linear m1(R1, T1, R1, T1), m2(R2, T2, R2, T2);
      C2A_TimeOfContactResult ra, rb;
      const double thr = 25.0;
      C2A_QueryContact(&m1, &m2, &ra, object1_tested, object2_tested, thr);
      rb.cont_l.push_front(ContactF());  // must be cleared by the call
      C2A_QueryContactOnly(&rb, R1, T1, object1_tested, R2, T2, object2_tested, thr);
      if (ra.num_contact != rb.num_contact || ra.cont_l.size() != rb.cont_l.size()) { only_bad++; continue; }
      std::list<ContactF>::iterator ia = ra.cont_l.begin(), ib = rb.cont_l.begin();
      for (; ia != ra.cont_l.end(); ++ia, ++ib)
        if (ia->TriangleID_A != ib->TriangleID_A || ia->TriangleID_B != ib->TriangleID_B || ia->Distance != ib->Distance) { only_bad++; break; }
    }
    printf("CONTACTONLY_MISMATCH %d\n", only_bad);
  }

  // Drive the CA loop from the host with C2A_TimeOfContactStep, the way C2A_QueryTimeOfContact does
  // (C2A/src/C2A.cpp:2005-2143), and compare with the one-call result.
  int step_bad = 0;
  for (int f = 0; f < nframes && f < 12; f++)
  {
    PQP_REAL R1[3][3], T1[3], R1e[3][3], T1e[3], R2[3][3], T2[3], R2e[3][3], T2e[3];
    t00[f].Rotation().Get_Value(R1); t00[f].Translation().Get_Value(T1); t01[f].Rotation().Get_Value(R1e); t01[f].Translation().Get_Value(T1e);
    t10[f].Rotation().Get_Value(R2); t10[f].Translation().Get_Value(T2); t11[f].Rotation().Get_Value(R2e); t11[f].Translation().Get_Value(T2e);
    CInterpMotion_Linear m1(R1, T1, R1e, T1e), m2(R2, T2, R2e, T2e);
    C2A_TimeOfContactResult res;
    Tri *const sa = object1_tested->tris, *const sb = object2_tested->tris;  // same seeds for both ways
    res.last_triA = sa; res.last_triB = sb;
    const PQP_REAL tol = 0.0001;
    PQP_REAL whole = C2A_QueryTimeOfContact(&m1, &m2, &res, object1_tested, object2_tested, tol, tol, 0);
    const bool whole_free = res.collisionfree; const int whole_ca = res.numCA; const PQP_REAL whole_dist = res.distance;

    CInterpMotion_Linear s1(R1, T1, R1e, T1e), s2(R2, T2, R2e, T2e);
    C2A_TimeOfContactResult sr;
    sr.last_triA = sa; sr.last_triB = sb;
    sr.num_bv_tests = sr.num_tri_tests = 0; sr.UpboundTOC = 1; sr.numCA = 0; sr.mint = 1;
    C2A_TimeOfContactStep(&s1, &s2, &sr, R1, T1, object1_tested, R2, T2, object2_tested, tol, tol);
    PQP_REAL dist = sr.distance, mint = sr.mint, lamda = 0, lastLamda = mint, toc = 0;
    sr.numCA = 1;
    bool is_free = false, done = false;
    int nItrs = 0;
    while (dist > tol)
    {
      if (++nItrs > 150) break;
      if (mint >= 1.0) { is_free = true; done = true; break; }
      if (mint < tol) break;
      lamda += mint;
      if (lamda >= 1.0) { is_free = true; done = true; break; }
      lastLamda = lamda;
      sr.numCA++;
      s1.integrate(lamda, R1, T1); s2.integrate(lamda, R2, T2);
      sr.UpboundTOC = 1.0 - lamda;
      C2A_TimeOfContactStep(&s1, &s2, &sr, R1, T1, object1_tested, R2, T2, object2_tested, tol, tol);
      dist = sr.distance; mint = sr.mint;
    }
    if (!done) { toc = lastLamda; if (toc >= 1 - tol) toc = 0; }
    if (is_free != whole_free || toc != whole || sr.numCA != whole_ca || dist != whole_dist) step_bad++;
  }
  printf("STEP_MISMATCH %d\n", step_bad);

  // C2A_Distance at each frame's start poses, the way PQP is used: the models' last_tri carries from call to call
  object1_tested->last_tri = object1_tested->tris; object2_tested->last_tri = object2_tested->tris;
  for (int f = 0; f < nframes && f < 24; f++)
  {
    PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3];
    t00[f].Rotation().Get_Value(R1); t00[f].Translation().Get_Value(T1);
    t10[f].Rotation().Get_Value(R2); t10[f].Translation().Get_Value(T2);
    C2A_DistanceResult dr;
    if (C2A_Distance(&dr, R1, T1, object1_tested, R2, T2, object2_tested, 0.0, 0.0) != PQP_OK) return 8;
    printf("D %a %d %d %d %d %a %a %a %a %a %a\n", dr.Distance(), dr.t1, dr.t2, dr.NumBVTests(), dr.NumTriTests(), dr.P1()[0], dr.P1()[1],
           dr.P1()[2], dr.P2()[0], dr.P2()[1], dr.P2()[2]);
  }

  // C2A_Collide, both overloads, at each frame's END poses (many interpenetrate): pair ids folded into a checksum
  object1_tested->last_tri = object1_tested->tris; object2_tested->last_tri = object2_tested->tris;
  {
    PQP_CollideResult cr;   // reused across calls like the reference's callers do
    for (int f = 0; f < nframes && f < 24; f++)
    {
      PQP_REAL R1[3][3], T1[3], R2[3][3], T2[3];
      t01[f].Rotation().Get_Value(R1); t01[f].Translation().Get_Value(T1);
      t11[f].Rotation().Get_Value(R2); t11[f].Translation().Get_Value(T2);
      if (C2A_Collide(&cr, R1, T1, object1_tested, R2, T2, object2_tested, C2A_ALL_CONTACTS) != PQP_OK) return 11;
      unsigned long long h = 1469598103934665603ull;
      for (int k = 0; k < cr.NumPairs(); k++) { h = (h ^ (unsigned)cr.Id1(k)) * 1099511628211ull; h = (h ^ (unsigned)cr.Id2(k)) * 1099511628211ull; }
      const int all = cr.NumPairs(), nbv = cr.NumBVTests(), ntri = cr.NumTriTests();
      if (C2A_Collide(&cr, R1, T1, object1_tested, R2, T2, object2_tested, C2A_FIRST_CONTACT) != PQP_OK) return 11;
      C2A_DistanceResult dr;
      if (C2A_Collide(&dr, R1, T1, object1_tested, R2, T2, object2_tested, 0.0, 0.0) != PQP_OK) return 11;
      printf("C %d %d %d %llu %d %d %a %d %d\n", all, nbv, ntri, h, cr.NumPairs(), cr.Colliding(), dr.Distance(), dr.t1, dr.t2);
    }
  }

  // pure translations through C2A_Solve (the reference's translation-only branch): optional 4th/5th arguments =
  // a pose file and a count; seeds are triangle 0 of each model, as the fixture was generated
  if (argc >= 6)
  {
    const int nt2 = atoi(argv[5]);
    FILE *ft = fopen(argv[4], "r");
    std::vector<double> tp(48 * nt2);
    for (int i = 0; i < 48 * nt2; i++) if (!ft || fscanf(ft, "%lf", &tp[i]) != 1) return 9;
    fclose(ft);
    for (int f = 0; f < nt2; f++)
    {
      Transform a0, a1, b0, b1, o0, o1;
      set_transform(a0, &tp[48 * f]); set_transform(a1, &tp[48 * f + 12]);
      set_transform(b0, &tp[48 * f + 24]); set_transform(b1, &tp[48 * f + 36]);
      C2A_TimeOfContactResult tr;
      tr.last_triA = object1_tested->tris; tr.last_triB = object2_tested->tris;
      PQP_REAL toc; int nItr, NTr;
      if (C2A_Solve(&a0, &a1, object1_tested, &b0, &b1, object2_tested, o0, o1, toc, nItr, NTr, 0.0, tr) != TOCFound) return 10;
      printf("T %d %a %a %d %d %d %d %d %d\n", tr.collisionfree ? 1 : 0, toc, tr.Distance(), nItr, tr.NumBVTests(), tr.NumTriTests(), NTr,
             (int)((C2A_Tri *)tr.last_triA - object1_tested->tris), (int)((C2A_Tri *)tr.last_triB - object2_tested->tris));
    }
  }
  delete[] cf;
  delete object1_tested;
  delete object2_tested;
  return 0;
}
