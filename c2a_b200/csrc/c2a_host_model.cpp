// Host-side RSS bounding-volume-hierarchy builder: the part of C2A_Model::EndModel() that feeds the
// CCD hot path.  It stays on the host (built once per mesh); its output is the flattened node/triangle
// layout of include/c2a_b200.h that c2a_b200_model_upload() moves to the GPU.
//
// Follows, with identical FP64 arithmetic order so that the tree is bit-identical to the reference's:
//   EndModel / C2A_BuildModel / build_recurse     C2A/src/C2A_PQP.cpp:331-417, C2A/src/C2A_Build.cpp:393-574
//   get_covariance_triverts, get_centroid_triverts, split_tris    C2A/src/C2A_Build.cpp:252-385
//   C2A_BV::FitToTris_Corner (RSS branch), ComputeAngularRadius    C2A/src/C2A_BV.cpp:354-635, 71-111
//   make_parent_relative                           C2A/src/C2A_Build.cpp:493-544
//   Meigen (PQP, un-vendored): cyclic Jacobi eigen-solver, the classical threshold-sweep form.
// Only the fields the traversals read are produced (R, Tr, l, r, R_loc, angularRadius, first_child, and the OBB
// half-dimensions d / centre To that C2A_Collide's overlap test reads, C2A_BV.cpp:398-418); corners and the
// uninitialised bookkeeping of the reference are not.
// Compile with -ffp-contract=off.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/c2a_b200.h"

namespace c2a_host {

struct Tri9 { double p[9]; int id; };

struct HostBvh
{
  std::vector<double> R, Tr, l, r, R_loc, ang;
  std::vector<double> obb_d, obb_To;  // [n][3] each: OBB half-dimensions and centre (parent-relative like Tr)
  std::vector<int32_t> first_child;
  std::vector<double> tris;      // permuted order
  std::vector<int32_t> tri_ids;  // original index of each permuted triangle
  std::vector<int32_t> tri_vidx; // vertex indices per permuted triangle (empty if not given)
  int depth = 0;
};

static inline double dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]); }
static inline double len3(const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
// r = M^T v, M row-major [3][3]
static inline void mtv(double r[3], const double M[3][3], const double v[3])
{
  r[0] = (M[0][0] * v[0] + M[1][0] * v[1] + M[2][0] * v[2]);
  r[1] = (M[0][1] * v[0] + M[1][1] * v[1] + M[2][1] * v[2]);
  r[2] = (M[0][2] * v[0] + M[1][2] * v[1] + M[2][2] * v[2]);
}
static inline void mv(double r[3], const double M[3][3], const double v[3])
{
  r[0] = (M[0][0] * v[0] + M[0][1] * v[1] + M[0][2] * v[2]);
  r[1] = (M[1][0] * v[0] + M[1][1] * v[1] + M[1][2] * v[2]);
  r[2] = (M[2][0] * v[0] + M[2][1] * v[1] + M[2][2] * v[2]);
}
static inline double max0(double a) { return a > 0 ? a : 0; }

// Symmetric 3x3 eigen-decomposition, cyclic Jacobi with the threshold schedule of the classical
// routine (first three sweeps rotate only above 0.2*off/9; tiny off-diagonals are zeroed after the
// fourth).  vout columns = eigenvectors.  a is destroyed.
static void jacobi3(double vout[3][3], double dout[3], double a[3][3])
{
  const int n = 3;
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, d[3], b[3], z[3];
  for (int p = 0; p < n; p++) { b[p] = d[p] = a[p][p]; z[p] = 0.0; }
  for (int sweep = 0; sweep < 50; sweep++)
  {
    double off = 0.0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += fabs(a[p][q]);
    if (off == 0.0)
    {
      memcpy(vout, v, sizeof(v));
      memcpy(dout, d, sizeof(d));
      return;
    }
    const double thresh = (sweep < 3) ? 0.2 * off / (n * n) : 0.0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++)
      {
        const double g = 100.0 * fabs(a[p][q]);
        if (sweep > 3 && fabs(d[p]) + g == fabs(d[p]) && fabs(d[q]) + g == fabs(d[q])) a[p][q] = 0.0;
        else if (fabs(a[p][q]) > thresh)
        {
          double h = d[q] - d[p], t;
          if (fabs(h) + g == fabs(h)) t = a[p][q] / h;
          else
          {
            const double theta = 0.5 * h / a[p][q];
            t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          const double c = 1.0 / sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
          h = t * a[p][q];
          z[p] -= h; z[q] += h; d[p] -= h; d[q] += h;
          a[p][q] = 0.0;
          auto rot = [&](double m[3][3], int i, int j, int k, int l) {
            const double g_ = m[i][j], h_ = m[k][l];
            m[i][j] = g_ - s * (h_ + g_ * tau);
            m[k][l] = h_ + s * (g_ - h_ * tau);
          };
          for (int j = 0; j < p; j++) rot(a, j, p, j, q);
          for (int j = p + 1; j < q; j++) rot(a, p, j, j, q);
          for (int j = q + 1; j < n; j++) rot(a, p, j, q, j);
          for (int j = 0; j < n; j++) rot(v, j, p, j, q);
        }
      }
    for (int p = 0; p < n; p++) { b[p] += z[p]; d[p] = b[p]; z[p] = 0.0; }
  }
  // not converged in 50 sweeps: the reference prints a warning and leaves its outputs unset; give
  // the current estimate instead
  memcpy(vout, v, sizeof(v));
  memcpy(dout, d, sizeof(d));
}

struct Builder
{
  std::vector<Tri9> &tris;
  HostBvh &out;
  int num_bvs = 0;
  std::vector<double> P;  // scratch: projected points

  Builder(std::vector<Tri9> &t, HostBvh &o) : tris(t), out(o) {}

  void covariance(double M[3][3], int first, int n)
  {
    double S1[3] = {0, 0, 0}, S2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < n; i++)
    {
      const double *p1 = tris[first + i].p, *p2 = p1 + 3, *p3 = p1 + 6;
      S1[0] += p1[0] + p2[0] + p3[0];
      S1[1] += p1[1] + p2[1] + p3[1];
      S1[2] += p1[2] + p2[2] + p3[2];
      S2[0][0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      S2[1][1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      S2[2][2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      S2[0][1] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      S2[0][2] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      S2[1][2] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
    }
    const double nn = (double)(3 * n);
    M[0][0] = S2[0][0] - S1[0] * S1[0] / nn;
    M[1][1] = S2[1][1] - S1[1] * S1[1] / nn;
    M[2][2] = S2[2][2] - S1[2] * S1[2] / nn;
    M[0][1] = S2[0][1] - S1[0] * S1[1] / nn;
    M[1][2] = S2[1][2] - S1[1] * S1[2] / nn;
    M[0][2] = S2[0][2] - S1[0] * S1[2] / nn;
    M[1][0] = M[0][1]; M[2][0] = M[0][2]; M[2][1] = M[1][2];
  }

  void centroid(double c[3], int first, int n)
  {
    c[0] = c[1] = c[2] = 0.0;
    for (int i = 0; i < n; i++)
    {
      const double *p1 = tris[first + i].p, *p2 = p1 + 3, *p3 = p1 + 6;
      c[0] += p1[0] + p2[0] + p3[0];
      c[1] += p1[1] + p2[1] + p3[1];
      c[2] += p1[2] + p2[2] + p3[2];
    }
    const double nn = (double)(3 * n);
    c[0] /= nn; c[1] /= nn; c[2] /= nn;
  }

  int split(int first, int n, const double a[3], double c)
  {
    int c1 = 0;
    for (int i = 0; i < n; i++)
    {
      const double *t = tris[first + i].p;
      double p[3] = {t[0], t[1], t[2]};
      p[0] = p[0] + t[3]; p[1] = p[1] + t[4]; p[2] = p[2] + t[5];
      p[0] = p[0] + t[6]; p[1] = p[1] + t[7]; p[2] = p[2] + t[8];
      double x = dot3(p, a);
      x /= 3.0;
      if (x <= c)
      {
        Tri9 tmp = tris[first + i];
        tris[first + i] = tris[first + c1];
        tris[first + c1] = tmp;
        c1++;
      }
    }
    if ((c1 == 0) || (c1 == n)) c1 = n / 2;
    return c1;
  }

  // RSS fit of node bn to tris [first, first+n) in orientation O (model frame)
  void fit(int bn, const double O[3][3], int first, int n)
  {
    // angular radius about the MODEL ORIGIN (the reference zeroes comRoot, quirk Q2)
    double ar = 0.0;
    for (int i = 0; i < n; i++)
      for (int v = 0; v < 3; v++)
      {
        const double *p = tris[first + i].p + 3 * v;
        const double pr[3] = {p[0] - 0.0, p[1] - 0.0, p[2] - 0.0};
        const double d = len3(pr);
        if (d > ar) ar = d;
      }
    out.ang[bn] = ar;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) { out.R[9 * bn + 3 * i + j] = O[i][j]; out.R_loc[9 * bn + 3 * i + j] = O[i][j]; }

    const int np = 3 * n;
    P.resize((size_t)3 * np);
    for (int i = 0; i < n; i++)
      for (int v = 0; v < 3; v++) mtv(&P[(size_t)3 * (3 * i + v)], O, tris[first + i].p + 3 * v);
    auto X = [&](int i) { return P[(size_t)3 * i]; };
    auto Y = [&](int i) { return P[(size_t)3 * i + 1]; };
    auto Z = [&](int i) { return P[(size_t)3 * i + 2]; };

    double minx, maxx, miny, maxy, minz, maxz;
    {
      // OBB fit (C2A_BV.cpp:398-418): extents of the projected points, centre back in the model frame
      minx = maxx = X(0); miny = maxy = Y(0); minz = maxz = Z(0);
      for (int i = 1; i < np; i++)
      {
        if (X(i) < minx) minx = X(i); else if (X(i) > maxx) maxx = X(i);
        if (Y(i) < miny) miny = Y(i); else if (Y(i) > maxy) maxy = Y(i);
        if (Z(i) < minz) minz = Z(i); else if (Z(i) > maxz) maxz = Z(i);
      }
      const double c[3] = {0.5 * (maxx + minx), 0.5 * (maxy + miny), 0.5 * (maxz + minz)};
      mv(&out.obb_To[3 * bn], O, c);
      out.obb_d[3 * bn] = 0.5 * (maxx - minx); out.obb_d[3 * bn + 1] = 0.5 * (maxy - miny); out.obb_d[3 * bn + 2] = 0.5 * (maxz - minz);
    }
    minz = maxz = Z(0);
    for (int i = 1; i < np; i++)
    {
      if (Z(i) < minz) minz = Z(i);
      else if (Z(i) > maxz) maxz = Z(i);
    }
    const double r = 0.5 * (maxz - minz), radsqr = r * r, cz = 0.5 * (maxz + minz);

    int minindex = 0, maxindex = 0;
    for (int i = 1; i < np; i++)
    {
      if (X(i) < X(minindex)) minindex = i;
      else if (X(i) > X(maxindex)) maxindex = i;
    }
    double x, y, dz;
    dz = Z(minindex) - cz; minx = X(minindex) + sqrt(max0(radsqr - dz * dz));
    dz = Z(maxindex) - cz; maxx = X(maxindex) - sqrt(max0(radsqr - dz * dz));
    for (int i = 0; i < np; i++)
      if (X(i) < minx)
      {
        dz = Z(i) - cz;
        x = X(i) + sqrt(max0(radsqr - dz * dz));
        if (x < minx) minx = x;
      }
    for (int i = 0; i < np; i++)
      if (X(i) > maxx)
      {
        dz = Z(i) - cz;
        x = X(i) - sqrt(max0(radsqr - dz * dz));
        if (x > maxx) maxx = x;
      }

    minindex = maxindex = 0;
    for (int i = 1; i < np; i++)
    {
      if (Y(i) < Y(minindex)) minindex = i;
      else if (Y(i) > Y(maxindex)) maxindex = i;
    }
    dz = Z(minindex) - cz; miny = Y(minindex) + sqrt(max0(radsqr - dz * dz));
    dz = Z(maxindex) - cz; maxy = Y(maxindex) - sqrt(max0(radsqr - dz * dz));
    for (int i = 0; i < np; i++)
      if (Y(i) < miny)
      {
        dz = Z(i) - cz;
        y = Y(i) + sqrt(max0(radsqr - dz * dz));
        if (y < miny) miny = y;
      }
    for (int i = 0; i < np; i++)
      if (Y(i) > maxy)
      {
        dz = Z(i) - cz;
        y = Y(i) - sqrt(max0(radsqr - dz * dz));
        if (y > maxy) maxy = y;
      }

    // corners: grow the rectangle where a point is outside both an x and a y bound
    const double a = sqrt(0.5);
    double dx, dy, u, t;
    for (int i = 0; i < np; i++)
    {
      if (X(i) > maxx)
      {
        if (Y(i) > maxy)
        {
          dx = X(i) - maxx; dy = Y(i) - maxy;
          u = dx * a + dy * a;
          t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - Z(i)) * (cz - Z(i));
          u = u - sqrt(max0(radsqr - t));
          if (u > 0) { maxx += u * a; maxy += u * a; }
        }
        else if (Y(i) < miny)
        {
          dx = X(i) - maxx; dy = Y(i) - miny;
          u = dx * a - dy * a;
          t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - Z(i)) * (cz - Z(i));
          u = u - sqrt(max0(radsqr - t));
          if (u > 0) { maxx += u * a; miny -= u * a; }
        }
      }
      else if (X(i) < minx)
      {
        if (Y(i) > maxy)
        {
          dx = X(i) - minx; dy = Y(i) - maxy;
          u = dy * a - dx * a;
          t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - Z(i)) * (cz - Z(i));
          u = u - sqrt(max0(radsqr - t));
          if (u > 0) { minx -= u * a; maxy += u * a; }
        }
        else if (Y(i) < miny)
        {
          dx = X(i) - minx; dy = Y(i) - miny;
          u = -dx * a - dy * a;
          t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - Z(i)) * (cz - Z(i));
          u = u - sqrt(max0(radsqr - t));
          if (u > 0) { minx -= u * a; miny -= u * a; }
        }
      }
    }

    const double corner[3] = {minx, miny, cz};
    mv(&out.Tr[3 * bn], O, corner);
    out.r[bn] = r;
    double l0 = maxx - minx, l1 = maxy - miny;
    if (l0 < 0) l0 = 0;
    if (l1 < 0) l1 = 0;
    out.l[2 * bn] = l0; out.l[2 * bn + 1] = l1;
  }

  void recurse(int bn, int first, int n, int depth)
  {
    if (depth > out.depth) out.depth = depth;
    double C[3][3], E[3][3], R[3][3], s[3];
    covariance(C, first, n);
    jacobi3(E, s, C);
    int mn, md, mx;
    if (s[0] > s[1]) { mx = 0; mn = 1; } else { mn = 0; mx = 1; }
    if (s[2] < s[mn]) { md = mn; mn = 2; }
    else if (s[2] > s[mx]) { md = mx; mx = 2; }
    else md = 2;
    for (int i = 0; i < 3; i++) { R[i][0] = E[i][mx]; R[i][1] = E[i][md]; }
    R[0][2] = E[1][mx] * E[2][md] - E[1][md] * E[2][mx];
    R[1][2] = E[0][md] * E[2][mx] - E[0][mx] * E[2][md];
    R[2][2] = E[0][mx] * E[1][md] - E[0][md] * E[1][mx];
    fit(bn, R, first, n);

    if (n == 1) out.first_child[bn] = -(first + 1);
    else
    {
      const int fc = num_bvs;
      out.first_child[bn] = fc;
      num_bvs += 2;
      const double axis[3] = {R[0][0], R[1][0], R[2][0]};
      double mean[3];
      centroid(mean, first, n);
      const double coord = dot3(axis, mean);
      const int n1 = split(first, n, axis, coord);
      recurse(fc, first, n1, depth + 1);
      recurse(fc + 1, first + n1, n - n1, depth + 1);
    }
  }

  // world-relative -> parent-relative (children first, then self)
  void parent_relative(int bn, const double pR[9], const double pT[3], const double pTo[3])
  {
    const int fc = out.first_child[bn];
    if (fc >= 0)
    {
      double myR[9], myT[3], myTo[3];
      memcpy(myR, &out.R[9 * bn], sizeof(myR));
      memcpy(myT, &out.Tr[3 * bn], sizeof(myT));
      memcpy(myTo, &out.obb_To[3 * bn], sizeof(myTo));
      parent_relative(fc, myR, myT, myTo);
      parent_relative(fc + 1, myR, myT, myTo);
    }
    {
      double *To = &out.obb_To[3 * bn];
      const double d[3] = {To[0] - pTo[0], To[1] - pTo[1], To[2] - pTo[2]};
      To[0] = (pR[0] * d[0] + pR[3] * d[1] + pR[6] * d[2]);
      To[1] = (pR[1] * d[0] + pR[4] * d[1] + pR[7] * d[2]);
      To[2] = (pR[2] * d[0] + pR[5] * d[1] + pR[8] * d[2]);
    }
    double *R = &out.R[9 * bn], *T = &out.Tr[3 * bn], Rpc[9], Tpc[3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Rpc[3 * i + j] = (pR[0 + i] * R[0 + j] + pR[3 + i] * R[3 + j] + pR[6 + i] * R[6 + j]);
    Tpc[0] = T[0] - pT[0]; Tpc[1] = T[1] - pT[1]; Tpc[2] = T[2] - pT[2];
    memcpy(R, Rpc, sizeof(Rpc));
    T[0] = (pR[0] * Tpc[0] + pR[3] * Tpc[1] + pR[6] * Tpc[2]);
    T[1] = (pR[1] * Tpc[0] + pR[4] * Tpc[1] + pR[7] * Tpc[2]);
    T[2] = (pR[2] * Tpc[0] + pR[5] * Tpc[1] + pR[8] * Tpc[2]);
  }
};

static void build(const double *tris9, int n, HostBvh &out)
{
  std::vector<Tri9> tris(n);
  for (int i = 0; i < n; i++) { memcpy(tris[i].p, tris9 + 9 * (size_t)i, sizeof(double) * 9); tris[i].id = i; }
  const int nb = 2 * n - 1;
  out.R.assign((size_t)9 * nb, 0); out.Tr.assign((size_t)3 * nb, 0); out.l.assign((size_t)2 * nb, 0);
  out.r.assign(nb, 0); out.obb_d.assign((size_t)3 * nb, 0); out.obb_To.assign((size_t)3 * nb, 0); out.R_loc.assign((size_t)9 * nb, 0); out.ang.assign(nb, 0); out.first_child.assign(nb, 0);
  Builder b(tris, out);
  b.num_bvs = 1;
  b.recurse(0, 0, n, 0);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z[3] = {0, 0, 0};
  b.parent_relative(0, I, Z, Z);
  out.tris.resize((size_t)9 * n);
  out.tri_ids.resize(n);
  for (int i = 0; i < n; i++) { memcpy(&out.tris[(size_t)9 * i], tris[i].p, sizeof(double) * 9); out.tri_ids[i] = tris[i].id; }
}

}  // namespace c2a_host

struct c2a_b200_host_bvh
{
  c2a_host::HostBvh b;
  int n_tris;
};

extern "C" {

int c2a_b200_bvh_build_indexed(const double *tris9, const int32_t *vidx3, int32_t n_tris, c2a_b200_host_bvh **out)
{
  if (!tris9 || !out || n_tris <= 0) return C2A_B200_ERR_ARG;
  c2a_b200_host_bvh *h = new c2a_b200_host_bvh();
  h->n_tris = n_tris;
  c2a_host::build(tris9, n_tris, h->b);
  if (vidx3)
  {
    h->b.tri_vidx.resize((size_t)3 * n_tris);
    for (int i = 0; i < n_tris; i++)
      for (int k = 0; k < 3; k++) h->b.tri_vidx[(size_t)3 * i + k] = vidx3[(size_t)3 * h->b.tri_ids[i] + k];
  }
  *out = h;
  return C2A_B200_OK;
}

int c2a_b200_bvh_build(const double *tris9, int32_t n_tris, c2a_b200_host_bvh **out)
{
  return c2a_b200_bvh_build_indexed(tris9, nullptr, n_tris, out);
}

int c2a_b200_bvh_view(const c2a_b200_host_bvh *h, c2a_b200_bvh *view, const int32_t **tri_ids, int32_t *depth)
{
  if (!h || !view) return C2A_B200_ERR_ARG;
  view->n_nodes = 2 * h->n_tris - 1;
  view->n_tris = h->n_tris;
  view->R = h->b.R.data(); view->Tr = h->b.Tr.data(); view->l = h->b.l.data(); view->r = h->b.r.data();
  view->R_loc = h->b.R_loc.data(); view->ang_radius = h->b.ang.data(); view->first_child = h->b.first_child.data();
  view->tris = h->b.tris.data();
  view->tri_vidx = h->b.tri_vidx.empty() ? nullptr : h->b.tri_vidx.data();
  view->obb_d = h->b.obb_d.data(); view->obb_To = h->b.obb_To.data();
  if (tri_ids) *tri_ids = h->b.tri_ids.data();
  if (depth) *depth = h->b.depth;
  return C2A_B200_OK;
}

int c2a_b200_bvh_free(c2a_b200_host_bvh *h)
{
  delete h;
  return C2A_B200_OK;
}

}  // extern "C"
