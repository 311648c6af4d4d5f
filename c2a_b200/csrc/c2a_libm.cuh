// sin / cos bit-identical to the platform libm the reference runs on.
//
// The reference integrates each object's pose with the C library's sin() and cos()
// (CInterpMotion::DeltaRt, /root/reference/C2A/src/InterpMotion.cpp:273-279).  CUDA's own sin/cos are
// accurate to 1-2 ulp but not identical to glibc's, and one ulp in a pose flips FP64 predicates in the
// controlled traversal (a different step sequence, a reported distance outside the 1e-9 contract).
// So the device evaluates sin/cos with the same algorithm as this image's libm: glibc 2.39, x86-64,
// the variant its ifunc selects on FMA-capable CPUs (table of sin/cos(k/128) in double-double plus
// short polynomials, IBM Accurate Mathematical Library lineage).  The fused-multiply-add placement
// below is the one that variant executes, so every intermediate rounds identically.
//
// Coverage: |x| < 105414350 (the range libm handles without its slow multi-precision reduction);
// beyond that CUDA's sin/cos are used (rotations that large do not occur: x = w*t/2 <= pi).
// Verified against math.sin/math.cos in tests/ (host mirror on CPU, device on the GPU box).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace c2a {

#define C2A_HD __host__ __device__ __forceinline__

static const double h_sincostab[440] = {
#include "c2a_sincostab.inc"
};
static __device__ const double d_sincostab[440] = {
#include "c2a_sincostab.inc"
};

#ifdef __CUDA_ARCH__
#define C2A_SCTAB(i) __ldg(&d_sincostab[(i)])
#define C2A_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define C2A_SCTAB(i) h_sincostab[(i)]
#define C2A_FMA(a, b, c) fma((a), (b), (c))
#endif

namespace lm {
constexpr double big = 0x1.8p+45;
constexpr double sn3 = -0x1.5555555555515p-3, sn5 = 0x1.11110e829872fp-7;
constexpr double cs2 = 0.5, cs4 = -0x1.5555555555535p-5, cs6 = 0x1.6c16bedd9e239p-10;
constexpr double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13,
                 s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26;
constexpr double hp0 = 0x1.921fb54442d18p+0, hp1 = 0x1.1a62633145c07p-54;
constexpr double toint = 0x1.8p+52, hpinv = 0x1.45f306dc9c883p-1;
constexpr double mp1 = 0x1.921fb58000000p+0, mp2 = -0x1.dde973c000000p-27, pp3 = -0x1.cb3b398000000p-55,
                 pp4 = -0x1.d747f23e32ed7p-83;
}  // namespace lm

C2A_HD uint64_t dbl_bits(double x)
{
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
C2A_HD double bits_dbl(uint64_t u)
{
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}
C2A_HD double copysign_bits(double mag, double sgn)
{
  return bits_dbl((dbl_bits(mag) & 0x7fffffffffffffffull) | (dbl_bits(sgn) & 0x8000000000000000ull));
}

// a + da with |a| < 0.126: odd polynomial with the low part folded in
C2A_HD double lm_taylor_sin(double a, double da)
{
  const double xx = a * a;
  double p = C2A_FMA(xx, lm::s5, lm::s4);
  p = C2A_FMA(xx, p, lm::s3);
  p = C2A_FMA(xx, p, lm::s2);
  p = C2A_FMA(xx, p, lm::s1);
  const double h = 0.5 * da;
  double t = C2A_FMA(p, a, -h);
  t = C2A_FMA(xx, t, da);
  return t + a;
}

// sin(xa + dx), xa = |x| in [0.126, 0.8555), dx the (sign-adjusted) low part; result >= 0
C2A_HD double lm_sin_tab(double xa, double dx)
{
  const double u = lm::big + xa;
  const double x = xa - (u - lm::big);
  const int k = (int)((uint32_t)dbl_bits(u)) * 4;
  const double xx = x * x;
  const double p = C2A_FMA(xx, lm::sn5, lm::sn3);
  const double s = x + C2A_FMA(x * xx, p, dx);
  double q = C2A_FMA(xx, lm::cs6, lm::cs4);
  q = C2A_FMA(xx, q, lm::cs2);
  const double c = C2A_FMA(x, dx, xx * q);
  const double sn = C2A_SCTAB(k), ssn = C2A_SCTAB(k + 1), cs = C2A_SCTAB(k + 2), ccs = C2A_SCTAB(k + 3);
  double cor = C2A_FMA(s, ccs, ssn);
  cor = C2A_FMA(-c, sn, cor);
  cor = C2A_FMA(s, cs, cor);
  return sn + cor;
}

// cos(xa + dx), xa = |x| < 0.8555
C2A_HD double lm_cos_tab(double xa, double dx)
{
  const double u = lm::big + xa;
  const double x = (xa - (u - lm::big)) + dx;
  const int k = (int)((uint32_t)dbl_bits(u)) * 4;
  const double xx = x * x;
  const double p = C2A_FMA(xx, lm::sn5, lm::sn3);
  const double s = C2A_FMA(x * xx, p, x);
  double q = C2A_FMA(xx, lm::cs6, lm::cs4);
  q = C2A_FMA(xx, q, lm::cs2);
  const double c = xx * q;
  const double sn = C2A_SCTAB(k), ssn = C2A_SCTAB(k + 1), cs = C2A_SCTAB(k + 2), ccs = C2A_SCTAB(k + 3);
  double cor = C2A_FMA(-s, ssn, ccs);
  cor = C2A_FMA(-c, cs, cor);
  cor = C2A_FMA(-s, sn, cor);
  return cs + cor;
}

// sin(a + da) for |a| < 0.8555 (libm's do_sin)
C2A_HD double lm_do_sin(double a, double da)
{
  const double aa = fabs(a);
  if (aa < 0.126) return lm_taylor_sin(a, da);
  const double dx = (a <= 0) ? -da : da;
  return copysign_bits(lm_sin_tab(aa, dx), a);
}
// cos(a + da) for |a| < 0.8555 (libm's do_cos)
C2A_HD double lm_do_cos(double a, double da)
{
  const double dx = (a < 0) ? -da : da;
  return lm_cos_tab(fabs(a), dx);
}

// x = n*(pi/2) + (a + da), |a| <= pi/4, for 2.426 < |x| < 105414350 (libm's reduce_sincos)
C2A_HD int lm_reduce(double x, double &a, double &da)
{
  const double t = C2A_FMA(x, lm::hpinv, lm::toint);
  const double xn = t - lm::toint;
  const int n = (int)((uint32_t)dbl_bits(t)) & 3;
  double y = C2A_FMA(-xn, lm::mp1, x);
  y = C2A_FMA(-xn, lm::mp2, y);
  const double t2 = C2A_FMA(-xn, lm::pp3, y);
  double d = C2A_FMA(-lm::pp3, xn, y - t2);
  a = C2A_FMA(-xn, lm::pp4, t2);
  const double e = C2A_FMA(-xn, lm::pp4, t2 - a);
  da = d + e;
  return n;
}

C2A_HD double lm_do_sincos(double a, double da, int n)
{
  double r = (n & 1) ? lm_do_cos(a, da) : lm_do_sin(a, da);
  return (n & 2) ? -r : r;
}

C2A_HD double libm_sin(double x)
{
  const uint32_t k = (uint32_t)(dbl_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;                       // |x| < 2^-26
  if (k < 0x3feb6000u) return lm_do_sin(x, 0.0);        // |x| < 0.855469
  if (k < 0x400368fdu)                                  // |x| < 2.426265
  {
    const double t = lm::hp0 - fabs(x);
    return copysign_bits(lm_do_cos(t, lm::hp1), x);
  }
  if (k < 0x419921FBu)
  {
    double a, da;
    const int n = lm_reduce(x, a, da);
    return lm_do_sincos(a, da, n);
  }
  return sin(x);
}

C2A_HD double libm_cos(double x)
{
  const uint32_t k = (uint32_t)(dbl_bits(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                      // |x| < 2^-27
  if (k < 0x3feb6000u) return lm_do_cos(x, 0.0);
  if (k < 0x400368fdu)
  {
    const double y = lm::hp0 - fabs(x);
    const double a = y + lm::hp1;
    const double da = (y - a) + lm::hp1;
    return lm_do_sin(a, da);
  }
  if (k < 0x419921FBu)
  {
    double a, da;
    const int n = lm_reduce(x, a, da);
    return lm_do_sincos(a, da, n + 1);
  }
  return cos(x);
}

}  // namespace c2a
