// Built-in forward-mode dynamics for the linearisation step (kernel K1).
//
// Dubins / unicycle car of the reference (`tests/dubins_car.py:48-90`): state (x, y, v, th),
// control (u_acc, u_turn), parameters p = (v_scale, w_scale, T), controls offset by +-eps
// (`:61-64`, zero derivative).  The reference divides by u2^2 and lets torch autograd
// differentiate that; its own fp64 values carry absolute noise ~1e-16/|u2|^3.
// Here the SAME function is evaluated without cancellation:  with D = T*a,
//     xp1 = x + v T (c P1 - s P2) + b T^2 (s P3 + c P4)
//     xp2 = y + v T (s P1 + c P2) + b T^2 (s P4 - c P3)
//     P1 = A, P2 = D B, P3 = D (C - B), P4 = A - B,
//     A = sin D/D, B = (1-cos D)/D^2, C = (D-sin D)/D^3, Dd = (D^2/2-1+cos D)/D^4, E = (D^3/6-D+sin D)/D^5
// (entire functions, power series for |D| <= 2).  oracle/dubins.py evaluates the same form (all five
// series summed directly); the two agree to ~1e-16.
#pragma once
#include <math.h>

namespace pmpc {

// A..E for j = 1..5:  sum_k (-1)^k x^k / (2k+j)!,  x = D^2.  Only the two highest-order series (Dd, E) are
// summed (Horner, 11 terms: |D| <= 2 -> truncation < 1e-19); the others follow from the exact relations
//     B = 1/2 - x Dd,   C = 1/6 - x E,   A = 1 - x C
// evaluated upwards, which is free of cancellation (x Dd <= 0.17 next to 1/2 for |D| <= 2).
__device__ __forceinline__ void dubins_entire(double D, double* A, double* B, double* C, double* Dd, double* E) {
  if (fabs(D) <= 2.0) {
    const double x = D * D;
    // reciprocal factorials 1/(2k+4)! and 1/(2k+5)!, k = 0..10, alternating sign folded in
    const double rd[11] = {1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600, -1.0 / 87178291200.0,
                           1.0 / 20922789888000.0, -1.0 / 6402373705728000.0, 1.0 / 2432902008176640000.0,
                           -1.0 / 1124000727777607680000.0, 1.0 / 620448401733239439360000.0};
    const double re[11] = {1.0 / 120, -1.0 / 5040, 1.0 / 362880, -1.0 / 39916800, 1.0 / 6227020800.0,
                           -1.0 / 1307674368000.0, 1.0 / 355687428096000.0, -1.0 / 121645100408832000.0,
                           1.0 / 51090942171709440000.0, -1.0 / 25852016738884976640000.0,
                           1.0 / 15511210043330985984000000.0};
    double d = rd[10], e = re[10];
#pragma unroll
    for (int k = 9; k >= 0; --k) {
      d = fma(d, x, rd[k]);
      e = fma(e, x, re[k]);
    }
    const double b = fma(-x, d, 0.5);
    const double c = fma(-x, e, 1.0 / 6.0);
    *A = fma(-x, c, 1.0); *B = b; *C = c; *Dd = d; *E = e;
  } else {
    double sD, cD;
    sincos(D, &sD, &cD);
    const double D2 = D * D;
    *A = sD / D;
    *B = (1.0 - cD) / D2;
    *C = (D - sD) / (D2 * D);
    *Dd = (0.5 * D2 - 1.0 + cD) / (D2 * D2);
    *E = (D2 * D / 6.0 - D + sD) / (D2 * D2 * D);
  }
}

// f[4], fx[16] row-major (fx[r*4+t] = d f_r / d x_t), fu[8] row-major (fu[r*2+t]).
__device__ __forceinline__ void dubins_f_fx_fu(const double* x, const double* u, double vs, double ws, double T,
                                               double eps, double* f, double* fx, double* fu) {
  double b = vs * u[0];
  double a = ws * (-u[1]);
  b += (b >= 0.0) ? eps : -eps;
  a += (a >= 0.0) ? eps : -eps;
  const double v0 = x[2], th = x[3];
  const double D = T * a;
  double A, B, C, Dd, E;
  dubins_entire(D, &A, &B, &C, &Dd, &E);
  double s, c;
  sincos(th, &s, &c);
  const double P1 = A, P2 = D * B, P3 = D * (C - B), P4 = A - B;
  const double dP1 = P3;
  const double dP2 = B + D * D * (2.0 * Dd - C);
  const double dP3 = (C - B) + D * D * (3.0 * E - 3.0 * Dd + C);
  const double dP4 = D * (2.0 * C - B - 2.0 * Dd);
  const double vT = v0 * T, bT2 = b * T * T;
  f[0] = x[0] + vT * (c * P1 - s * P2) + bT2 * (s * P3 + c * P4);
  f[1] = x[1] + vT * (s * P1 + c * P2) + bT2 * (s * P4 - c * P3);
  f[2] = v0 + T * b;
  f[3] = T * a + th;
#pragma unroll
  for (int i = 0; i < 16; ++i) fx[i] = 0.0;
  fx[0] = 1.0; fx[5] = 1.0; fx[10] = 1.0; fx[15] = 1.0;
  fx[2] = T * (c * P1 - s * P2);
  fx[6] = T * (s * P1 + c * P2);
  fx[3] = vT * (-s * P1 - c * P2) + bT2 * (c * P3 - s * P4);
  fx[7] = vT * (c * P1 - s * P2) + bT2 * (c * P4 + s * P3);
  const double d1_b = T * T * (s * P3 + c * P4);
  const double d2_b = T * T * (s * P4 - c * P3);
  const double d1_a = T * (vT * (c * dP1 - s * dP2) + bT2 * (s * dP3 + c * dP4));
  const double d2_a = T * (vT * (s * dP1 + c * dP2) + bT2 * (s * dP4 - c * dP3));
  fu[0] = vs * d1_b; fu[1] = -ws * d1_a;
  fu[2] = vs * d2_b; fu[3] = -ws * d2_a;
  fu[4] = vs * T;    fu[5] = 0.0;
  fu[6] = 0.0;       fu[7] = -ws * T;
}

// The entries of the sparse Dubins stage (fused.inl: lin_sparse) that need neither sin/cos nor the series:
// b20 = fu[4], b31 = fu[7], and the affine terms of the speed / heading rows  c2 = f2 - b20 u0 [- v0],
// c3 = f3 - b31 u1 [- th]  (the bracketed part for stages j >= 1 only), evaluated exactly as dubins_f_fx_fu and
// lin_sparse evaluate them, so that the forward sweep can rebuild them instead of reading them from HBM.
__device__ __forceinline__ void dubins_cheap_entries(const double* x, const double* u, double vs, double ws, double T, double eps,
                                                     bool has_a, double* b20, double* b31, double* c2, double* c3) {
  double b = vs * u[0];
  double a = ws * (-u[1]);
  b += (b >= 0.0) ? eps : -eps;
  a += (a >= 0.0) ? eps : -eps;
  const double f2 = x[2] + T * b, f3 = T * a + x[3];
  *b20 = vs * T;
  *b31 = -ws * T;
  double v2 = f2 - (*b20) * u[0], v3 = f3 - (*b31) * u[1];
  if (has_a) {
    v2 -= x[2];
    v3 -= x[3];
  }
  *c2 = v2;
  *c3 = v3;
}

}  // namespace pmpc
