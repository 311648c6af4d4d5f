// Device restatement of the reference's linear interpolated motion (sm_100a, FP64, -fmad=false).
//
//   (motion constants: host side, c2a_host_motion.h <- CInterpMotion ctor + velocity + LinearAngularVelocity
//                         C2A/src/InterpMotion.cpp:148-168, 486-491, 228-270)
//   motion_pose        <- CInterpMotion_Linear::integrate, AbsoluteRt, DeltaRt   :516-568, 273-287
//   motion_bound_bv    <- CInterpMotion_Linear::computeTOC_MotionBound           :831-881
//   motion_bound_leaf  <- CInterpMotion_Linear::computeTOC(d, r1, S)             :746-828
//   quat_from_matrix   <- Matrix3x3::Quaternion_          C2A/LinearMath.h:759-793
//   matrix_from_quat   <- Matrix3x3::Set_Value(Quaternion) C2A/LinearMath.h:809-831
//   quat_mul           <- Quaternion operator%             C2A/LinearMath.h:1018-1025
#pragma once
#include "c2a_geom.cuh"
#include "c2a_libm.cuh"

namespace c2a {

// Per-object motion constants.  qs is the start rotation's quaternion (x,y,z,w): the reference
// re-derives it from the start matrix on every integrate() (InterpMotion.cpp:284); the value is the
// same every time, so it is computed once.
struct Motion
{
  double cv[3];    // cv: linear velocity of the origin (T_end - T_start)
  double axis[3];  // m_axis (unit)
  double w;        // m_angVel
  double qs[4];
  double Ts[3];
};

C2A_HD void quat_from_matrix(double q[4], const double val[9])
{
  const double trace = val[0] + val[4] + val[8];
  if (trace > 0.0)
  {
    double s = sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (val[7] - val[5]) * s;
    q[1] = (val[2] - val[6]) * s;
    q[2] = (val[3] - val[1]) * s;
  }
  else
  {
    // i = index of the largest diagonal entry, (i,j,k) cyclic
    const int i = val[0] < val[4] ? (val[4] < val[8] ? 2 : 1) : (val[0] < val[8] ? 2 : 0);
    double dii, djj, dkk, kj, jk, ji, ij, ki, ik;
    if (i == 0) { dii = val[0]; djj = val[4]; dkk = val[8]; kj = val[7]; jk = val[5]; ji = val[3]; ij = val[1]; ki = val[6]; ik = val[2]; }
    else if (i == 1) { dii = val[4]; djj = val[8]; dkk = val[0]; kj = val[2]; jk = val[6]; ji = val[7]; ij = val[5]; ki = val[1]; ik = val[3]; }
    else { dii = val[8]; djj = val[0]; dkk = val[4]; kj = val[3]; jk = val[1]; ji = val[2]; ij = val[6]; ki = val[5]; ik = val[7]; }
    double s = sqrt(dii - djj - dkk + 1.0);
    const double qi = s * 0.5;
    s = 0.5 / s;
    const double qw = (kj - jk) * s, qj = (ji + ij) * s, qk = (ki + ik) * s;
    q[3] = qw;
    if (i == 0) { q[0] = qi; q[1] = qj; q[2] = qk; }
    else if (i == 1) { q[1] = qi; q[2] = qj; q[0] = qk; }
    else { q[2] = qi; q[0] = qj; q[1] = qk; }
  }
}

C2A_HD void quat_mul(double r[4], const double a[4], const double b[4])
{
  r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

C2A_DEV void matrix_from_quat(double v[9], const double q[4])
{
  const double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const double s = 2.0 / d;
  const double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
  const double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  const double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
  const double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  v[0] = 1.0 - (yy + zz); v[1] = xy - wz; v[2] = xz + wy;
  v[3] = xy + wz; v[4] = 1.0 - (xx + zz); v[5] = yz - wx;
  v[6] = xz - wy; v[7] = yz + wx; v[8] = 1.0 - (xx + yy);
}

// Motion record of one object (24 doubles), produced on the host by c2a_b200_motions_from_poses():
//   R0(9) T0(3) cv(3) axis(3) w qs(4) pad(1)
// The constants come from the host because LinearAngularVelocity needs acos(), and the reference's
// values are whatever the host libm returns (InterpMotion.cpp:250-252).
constexpr int MOTION_DOUBLES = 24;
C2A_DEV void motion_load(Motion &m, const double *rec)
{
#pragma unroll
  for (int i = 0; i < 3; i++) { m.Ts[i] = __ldg(rec + 9 + i); m.cv[i] = __ldg(rec + 12 + i); m.axis[i] = __ldg(rec + 15 + i); }
  m.w = __ldg(rec + 18);
#pragma unroll
  for (int i = 0; i < 4; i++) m.qs[i] = __ldg(rec + 19 + i);
}

// pose at time t (clamped to <= 1): T = T_s + cv*t, R from q_s (x) (axis sin(wt/2), cos(wt/2))
C2A_DEV void motion_pose(const Motion &m, double t_in, double R[9], double T[3])
{
  double dt = t_in;
  if (dt > 1) dt = 1;
  T[0] = m.Ts[0] + dt * m.cv[0];
  T[1] = m.Ts[1] + dt * m.cv[1];
  T[2] = m.Ts[2] + dt * m.cv[2];
  const double ang = 0.5 * m.w * dt;
  const double sn = libm_sin(ang), cs = libm_cos(ang);  // bit-identical to the host libm, see c2a_libm.cuh
  const double drt[4] = {sn * m.axis[0], sn * m.axis[1], sn * m.axis[2], cs};
  double q[4];
  quat_mul(q, m.qs, drt);
  matrix_from_quat(R, q);
}

// The reference normalises the direction inside each bound (Vnormalize, InterpMotion.cpp:837 / :752) and
// calls the second object's bound with S2 = S1 * -1.  Squares are sign-blind and negation is exact, so
// normalising S2 gives exactly -normalise(S1): the callers normalise once and pass the negated unit
// vector to the second bound (one FP64 sqrt + division less per child test, same bits).
C2A_DEV double motion_bound_bv_unit(const Motion &m, double ang_radius, const double N[3])
{
  double cross[3];
  v_cross(cross, m.axis, N);
  const double w_max = (ang_radius)*v_len(cross) * m.w;
  double v_max = v_dot(m.cv, N);
  if (v_max < 0) v_max = 0;
  double path_max = v_max + w_max;
  if (path_max <= 0) path_max = 1e-30;
  return path_max;
}
C2A_DEV double motion_bound_leaf_unit(const Motion &m, double ang_radius, const double S[3])
{
  double v_max, w_max;
  if (m.w == 0)
  {
    w_max = 0;
    v_max = v_dot(m.cv, S);
    if (v_max < 0) v_max = 0;
  }
  else
  {
    double cwc[3] = {m.axis[0], m.axis[1], m.axis[2]}, cross[3];
    cwc[0] *= m.w; cwc[1] *= m.w; cwc[2] *= m.w;
    v_cross(cross, cwc, S);
    w_max = ang_radius * v_len(cross);
    v_max = v_dot(m.cv, S);
    if (v_max < 0) v_max = 0;
  }
  double path_max = w_max + v_max;
  if (path_max == 0) path_max = 1e-30;
  return path_max;
}

// directional motion bound of a BV; normalises N in place like the reference
C2A_DEV double motion_bound_bv(const Motion &m, double ang_radius, double N[3])
{
  double cross[3];
  v_normalize(N);
  v_cross(cross, m.axis, N);
  const double w_max = (ang_radius)*v_len(cross) * m.w;
  double v_max = v_dot(m.cv, N);
  if (v_max < 0) v_max = 0;
  double path_max = v_max + w_max;
  if (path_max <= 0) path_max = 1e-30;
  return path_max;
}

// directional motion bound at a leaf; normalises S in place like the reference
C2A_DEV double motion_bound_leaf(const Motion &m, double ang_radius, double S[3])
{
  double v_max, w_max;
  v_normalize(S);
  if (m.w == 0)
  {
    w_max = 0;
    v_max = v_dot(m.cv, S);
    if (v_max < 0) v_max = 0;
  }
  else
  {
    double cwc[3] = {m.axis[0], m.axis[1], m.axis[2]}, cross[3];
    cwc[0] *= m.w; cwc[1] *= m.w; cwc[2] *= m.w;
    v_cross(cross, cwc, S);
    w_max = ang_radius * v_len(cross);
    v_max = v_dot(m.cv, S);
    if (v_max < 0) v_max = 0;
  }
  double path_max = w_max + v_max;
  if (path_max == 0) path_max = 1e-30;
  return path_max;
}

}  // namespace c2a
