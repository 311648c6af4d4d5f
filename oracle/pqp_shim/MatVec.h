// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's MatVec.h.
//
// PQP (GammaUNC/PQP, v1.3 lineage, unpinned: /root/reference/README.md:11) is not
// vendored by the reference and is absent from this image, so the reference's
// C2A sources are compiled against this restatement.  Conventions are the ones
// the reference's call sites imply (C2A/src/C2A_PQP.cpp:296-299,934-941):
// row-major M[r][c], MxM = A*B, MTxM = A^T*B, MxV = A*v, MTxV = A^T*v,
// MxVpV = A*v + t.  Every sum is evaluated left to right with no fused
// multiply-add (build with -ffp-contract=off) because the parity contract is
// bit-level on the traversal's FP64 predicates.
#ifndef PQP_SHIM_MATVEC_H
#define PQP_SHIM_MATVEC_H

#include <math.h>
#include <stdio.h>
#include "PQP_Compile.h"

#ifndef M_PI
const PQP_REAL M_PI = (PQP_REAL)3.14159265359;
#endif

#ifndef myfabs
#define myfabs(x) ((x < 0) ? -x : x)
#endif

inline void Midentity(PQP_REAL M[3][3])
{
  M[0][0] = M[1][1] = M[2][2] = 1.0;
  M[0][1] = M[1][2] = M[2][0] = 0.0;
  M[0][2] = M[1][0] = M[2][1] = 0.0;
}

inline void Videntity(PQP_REAL T[3]) { T[0] = T[1] = T[2] = 0.0; }

inline void McM(PQP_REAL Mr[3][3], const PQP_REAL M[3][3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Mr[i][j] = M[i][j];
}

inline void MTcM(PQP_REAL Mr[3][3], const PQP_REAL M[3][3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Mr[i][j] = M[j][i];
}

inline void VcV(PQP_REAL Vr[3], const PQP_REAL V[3])
{
  Vr[0] = V[0]; Vr[1] = V[1]; Vr[2] = V[2];
}

inline void McolcV(PQP_REAL Vr[3], const PQP_REAL M[3][3], int c)
{
  Vr[0] = M[0][c]; Vr[1] = M[1][c]; Vr[2] = M[2][c];
}

inline void McolcMcol(PQP_REAL Mr[3][3], int cr, const PQP_REAL M[3][3], int c)
{
  Mr[0][cr] = M[0][c]; Mr[1][cr] = M[1][c]; Mr[2][cr] = M[2][c];
}

inline void MxMpV(PQP_REAL Mr[3][3], const PQP_REAL M1[3][3], const PQP_REAL M2[3][3],
                  const PQP_REAL T[3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Mr[i][j] = (M1[i][0] * M2[0][j] + M1[i][1] * M2[1][j] + M1[i][2] * M2[2][j] + T[i]);
}

inline void MxM(PQP_REAL Mr[3][3], const PQP_REAL M1[3][3], const PQP_REAL M2[3][3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Mr[i][j] = (M1[i][0] * M2[0][j] + M1[i][1] * M2[1][j] + M1[i][2] * M2[2][j]);
}

inline void MxMT(PQP_REAL Mr[3][3], const PQP_REAL M1[3][3], const PQP_REAL M2[3][3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Mr[i][j] = (M1[i][0] * M2[j][0] + M1[i][1] * M2[j][1] + M1[i][2] * M2[j][2]);
}

inline void MTxM(PQP_REAL Mr[3][3], const PQP_REAL M1[3][3], const PQP_REAL M2[3][3])
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      Mr[i][j] = (M1[0][i] * M2[0][j] + M1[1][i] * M2[1][j] + M1[2][i] * M2[2][j]);
}

inline void MxV(PQP_REAL Vr[3], const PQP_REAL M1[3][3], const PQP_REAL V1[3])
{
  Vr[0] = (M1[0][0] * V1[0] + M1[0][1] * V1[1] + M1[0][2] * V1[2]);
  Vr[1] = (M1[1][0] * V1[0] + M1[1][1] * V1[1] + M1[1][2] * V1[2]);
  Vr[2] = (M1[2][0] * V1[0] + M1[2][1] * V1[1] + M1[2][2] * V1[2]);
}

inline void MxVpV(PQP_REAL Vr[3], const PQP_REAL M1[3][3], const PQP_REAL V1[3],
                  const PQP_REAL V2[3])
{
  Vr[0] = (M1[0][0] * V1[0] + M1[0][1] * V1[1] + M1[0][2] * V1[2] + V2[0]);
  Vr[1] = (M1[1][0] * V1[0] + M1[1][1] * V1[1] + M1[1][2] * V1[2] + V2[1]);
  Vr[2] = (M1[2][0] * V1[0] + M1[2][1] * V1[1] + M1[2][2] * V1[2] + V2[2]);
}

inline void sMxVpV(PQP_REAL Vr[3], PQP_REAL s1, const PQP_REAL M1[3][3], const PQP_REAL V1[3],
                   const PQP_REAL V2[3])
{
  Vr[0] = s1 * (M1[0][0] * V1[0] + M1[0][1] * V1[1] + M1[0][2] * V1[2]) + V2[0];
  Vr[1] = s1 * (M1[1][0] * V1[0] + M1[1][1] * V1[1] + M1[1][2] * V1[2]) + V2[1];
  Vr[2] = s1 * (M1[2][0] * V1[0] + M1[2][1] * V1[1] + M1[2][2] * V1[2]) + V2[2];
}

inline void MTxV(PQP_REAL Vr[3], const PQP_REAL M1[3][3], const PQP_REAL V1[3])
{
  Vr[0] = (M1[0][0] * V1[0] + M1[1][0] * V1[1] + M1[2][0] * V1[2]);
  Vr[1] = (M1[0][1] * V1[0] + M1[1][1] * V1[1] + M1[2][1] * V1[2]);
  Vr[2] = (M1[0][2] * V1[0] + M1[1][2] * V1[1] + M1[2][2] * V1[2]);
}

inline void sMTxV(PQP_REAL Vr[3], PQP_REAL s1, const PQP_REAL M1[3][3], const PQP_REAL V1[3])
{
  Vr[0] = s1 * (M1[0][0] * V1[0] + M1[1][0] * V1[1] + M1[2][0] * V1[2]);
  Vr[1] = s1 * (M1[0][1] * V1[0] + M1[1][1] * V1[1] + M1[2][1] * V1[2]);
  Vr[2] = s1 * (M1[0][2] * V1[0] + M1[1][2] * V1[1] + M1[2][2] * V1[2]);
}

inline void sMxV(PQP_REAL Vr[3], PQP_REAL s1, const PQP_REAL M1[3][3], const PQP_REAL V1[3])
{
  Vr[0] = s1 * (M1[0][0] * V1[0] + M1[0][1] * V1[1] + M1[0][2] * V1[2]);
  Vr[1] = s1 * (M1[1][0] * V1[0] + M1[1][1] * V1[1] + M1[1][2] * V1[2]);
  Vr[2] = s1 * (M1[2][0] * V1[0] + M1[2][1] * V1[1] + M1[2][2] * V1[2]);
}

inline void VmV(PQP_REAL Vr[3], const PQP_REAL V1[3], const PQP_REAL V2[3])
{
  Vr[0] = V1[0] - V2[0]; Vr[1] = V1[1] - V2[1]; Vr[2] = V1[2] - V2[2];
}

inline void VpV(PQP_REAL Vr[3], const PQP_REAL V1[3], const PQP_REAL V2[3])
{
  Vr[0] = V1[0] + V2[0]; Vr[1] = V1[1] + V2[1]; Vr[2] = V1[2] + V2[2];
}

inline void VpVxS(PQP_REAL Vr[3], const PQP_REAL V1[3], const PQP_REAL V2[3], PQP_REAL s)
{
  Vr[0] = V1[0] + V2[0] * s; Vr[1] = V1[1] + V2[1] * s; Vr[2] = V1[2] + V2[2] * s;
}

inline void VcrossV(PQP_REAL Vr[3], const PQP_REAL V1[3], const PQP_REAL V2[3])
{
  Vr[0] = V1[1] * V2[2] - V1[2] * V2[1];
  Vr[1] = V1[2] * V2[0] - V1[0] * V2[2];
  Vr[2] = V1[0] * V2[1] - V1[1] * V2[0];
}

inline PQP_REAL VdotV(const PQP_REAL V1[3], const PQP_REAL V2[3])
{
  return (V1[0] * V2[0] + V1[1] * V2[1] + V1[2] * V2[2]);
}

inline PQP_REAL VdistV2(const PQP_REAL V1[3], const PQP_REAL V2[3])
{
  return ((V1[0] - V2[0]) * (V1[0] - V2[0]) + (V1[1] - V2[1]) * (V1[1] - V2[1]) +
          (V1[2] - V2[2]) * (V1[2] - V2[2]));
}

inline void VxS(PQP_REAL Vr[3], const PQP_REAL V[3], PQP_REAL s)
{
  Vr[0] = V[0] * s; Vr[1] = V[1] * s; Vr[2] = V[2] * s;
}

inline PQP_REAL Vlength(const PQP_REAL V[3])
{
  return sqrt(V[0] * V[0] + V[1] * V[1] + V[2] * V[2]);
}

inline void Vnormalize(PQP_REAL V[3])
{
  PQP_REAL d = (PQP_REAL)1.0 / sqrt(V[0] * V[0] + V[1] * V[1] + V[2] * V[2]);
  V[0] *= d; V[1] *= d; V[2] *= d;
}

inline void MVtoOGL(double oglm[16], const PQP_REAL R[3][3], const PQP_REAL T[3])
{
  oglm[0] = R[0][0]; oglm[1] = R[1][0]; oglm[2] = R[2][0]; oglm[3] = 0.0;
  oglm[4] = R[0][1]; oglm[5] = R[1][1]; oglm[6] = R[2][1]; oglm[7] = 0.0;
  oglm[8] = R[0][2]; oglm[9] = R[1][2]; oglm[10] = R[2][2]; oglm[11] = 0.0;
  oglm[12] = T[0]; oglm[13] = T[1]; oglm[14] = T[2]; oglm[15] = 1.0;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi sweeps (the classical
// published algorithm: Jacobi 1846 / Rutishauser 1966, threshold variant).
// vout columns are eigenvectors, dout the eigenvalues; a is destroyed.
// Called once per BVH node by the reference's builder (C2A/src/C2A_Build.cpp:416).
// The product's own host builder (c2a_b200/csrc/c2a_host_model.cpp) restates the
// same sweeps so both sides build the identical tree.
inline void Meigen(PQP_REAL vout[3][3], PQP_REAL dout[3], PQP_REAL a[3][3])
{
  const int n = 3;
  PQP_REAL v[3][3], d[3], b[3], z[3];
  Midentity(v);
  for (int p = 0; p < n; p++) { b[p] = d[p] = a[p][p]; z[p] = 0.0; }

  for (int sweep = 0; sweep < 50; sweep++)
  {
    PQP_REAL off = 0.0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += fabs(a[p][q]);
    if (off == 0.0) { McM(vout, v); VcV(dout, d); return; }

    PQP_REAL thresh = (sweep < 3) ? (PQP_REAL)0.2 * off / (n * n) : (PQP_REAL)0.0;

    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++)
      {
        PQP_REAL g = (PQP_REAL)100.0 * fabs(a[p][q]);
        if (sweep > 3 && fabs(d[p]) + g == fabs(d[p]) && fabs(d[q]) + g == fabs(d[q]))
        {
          a[p][q] = 0.0;
        }
        else if (fabs(a[p][q]) > thresh)
        {
          PQP_REAL h = d[q] - d[p], t;
          if (fabs(h) + g == fabs(h)) t = a[p][q] / h;
          else
          {
            PQP_REAL theta = (PQP_REAL)0.5 * h / a[p][q];
            t = (PQP_REAL)(1.0 / (fabs(theta) + sqrt(1.0 + theta * theta)));
            if (theta < 0.0) t = -t;
          }
          PQP_REAL c = (PQP_REAL)1.0 / sqrt(1 + t * t);
          PQP_REAL s = t * c;
          PQP_REAL tau = s / ((PQP_REAL)1.0 + c);
          h = t * a[p][q];
          z[p] -= h; z[q] += h; d[p] -= h; d[q] += h;
          a[p][q] = 0.0;
#define PQP_SHIM_ROT(m, i, j, k, l)                       \
  { PQP_REAL g_ = m[i][j], h_ = m[k][l];                  \
    m[i][j] = g_ - s * (h_ + g_ * tau);                   \
    m[k][l] = h_ + s * (g_ - h_ * tau); }
          for (int j = 0; j < p; j++) PQP_SHIM_ROT(a, j, p, j, q)
          for (int j = p + 1; j < q; j++) PQP_SHIM_ROT(a, p, j, j, q)
          for (int j = q + 1; j < n; j++) PQP_SHIM_ROT(a, p, j, q, j)
          for (int j = 0; j < n; j++) PQP_SHIM_ROT(v, j, p, j, q)
#undef PQP_SHIM_ROT
        }
      }
    for (int p = 0; p < n; p++) { b[p] += z[p]; d[p] = b[p]; z[p] = 0.0; }
  }
  fprintf(stderr, "eigen: too many iterations in Jacobi transform.\n");
}

#endif
