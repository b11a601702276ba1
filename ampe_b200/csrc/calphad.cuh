// Per-cell CALPHAD binary two-phase thermodynamics and the KKS Newton solve, in
// registers.  Replaces the Thermo4PFM calls AMPE makes from
//   CALPHADequilibriumPhaseConcentrationsStrategy.cc:385-387  (computePhaseConcentrations)
//   CALPHADFreeEnergyStrategyBinary.cc:321-323, 676-680       (computeFreeEnergy, computeDerivFreeEnergy)
//   MobilityCompositionDiffusionStrategy.cc:57-69             (second derivative)
// and the in-tree CALPHADMobility.{h,cc} formulas.  Model equations:
// doc/latex/manual/appendix.tex:517-599.
#pragma once
#include "fastmath.cuh"
#include "params.h"
#include "pointwise.cuh"

namespace ampe {

// xlogx family with the C2 quadratic extension below 1e-8
#define AMPE_SMALLX 1.0e-8
AMPE_DEV double xlogx(double x, double log_smallx)
{
   if (x > AMPE_SMALLX) return x * log(x);
   const double inv = 1. / AMPE_SMALLX;
   return AMPE_SMALLX * log_smallx + (x - AMPE_SMALLX) * (1. + log_smallx) +
          0.5 * (x * x * inv - 2.0 * x + AMPE_SMALLX);
}
AMPE_DEV double xlogx_deriv(double x, double log_smallx)
{
   if (x > AMPE_SMALLX) return log(x) + 1.0;
   return (1. + log_smallx) + (x - AMPE_SMALLX) * (1. / AMPE_SMALLX);
}
AMPE_DEV double xlogx_deriv2(double x)
{
   if (x > AMPE_SMALLX) return 1. / x;
   return 1. / AMPE_SMALLX;
}
// log(1e-8) as glibc rounds it (host-evaluated constant, same value the CPU side uses)
#define AMPE_LOG_SMALLX (-18.420680743952367)
// the extension branch of xlogx_deriv (x <= 1e-8)
AMPE_DEV double xlogx_deriv_ext(double x) { return (1. + AMPE_LOG_SMALLX) + (x - AMPE_SMALLX) * (1. / AMPE_SMALLX); }
#ifdef AMPE_KKS_LIBM_LOG
#define AMPE_KKS_LOG(x) log(x)
#define AMPE_KKS_RCP(x) (1.0 / (x))
#else
#define AMPE_KKS_LOG(x) log_fast(x)
#define AMPE_KKS_RCP(x) rcp_fast(x)
#endif
// xlogx_deriv2 with the comparison already made: 1/x above the extension, 1e8 below
AMPE_DEV double xlogx_deriv2_fast(double x, bool above) { return above ? AMPE_KKS_RCP(above ? x : 1.0) : 1. / AMPE_SMALLX; }

AMPE_DEV double fmix(const double* L, double c)
{
   const double t = 2.0 * c - 1.0;
   return c * (1.0 - c) * (L[0] + L[1] * t + L[2] * t * t + L[3] * t * t * t);
}
AMPE_DEV double fmix_deriv(const double* L, double c)
{
   const double t = 2.0 * c - 1.0;
   const double cc = c * (1. - c);
   return (1.0 - 2.0 * c) * (L[0] + L[1] * t + L[2] * t * t + L[3] * t * t * t) +
          cc * (2.0 * L[1] + 4.0 * L[2] * t + 6.0 * L[3] * t * t);
}
AMPE_DEV double fmix_deriv2(const double* L, double c)
{
   const double t = 2.0 * c - 1.0;
   const double cc = c * (1. - c);
   return -2.0 * (L[0] + L[1] * t + L[2] * t * t + L[3] * t * t * t) +
          2.0 * (1.0 - 2.0 * c) * (2.0 * L[1] + 4.0 * L[2] * t + 6.0 * L[3] * t * t) +
          cc * (8.0 * L[2] + 24.0 * L[3] * t);
}

AMPE_DEV double calphad_f(const CalphadT& t, double c, int pi)
{
   return c * t.fA[pi] + (1.0 - c) * t.fB[pi] + fmix(t.L[pi], c) +
          t.RT * (xlogx(c, AMPE_LOG_SMALLX) + xlogx(1.0 - c, AMPE_LOG_SMALLX));
}
AMPE_DEV double calphad_mu(const CalphadT& t, double c, int pi)
{
   return (t.fA[pi] - t.fB[pi]) + fmix_deriv(t.L[pi], c) +
          t.RT * (xlogx_deriv(c, AMPE_LOG_SMALLX) - xlogx_deriv(1.0 - c, AMPE_LOG_SMALLX));
}
AMPE_DEV double calphad_d2f(const CalphadT& t, double c, int pi)
{
   return fmix_deriv2(t.L[pi], c) + t.RT * (xlogx_deriv2(c) + xlogx_deriv2(1.0 - c));
}

// Redlich-Kister pieces of one phase at concentration c, Horner / fma form:
//   P = sum L_k t^k, dP = dP/dc = 2 L_1 + 4 L_2 t + 6 L_3 t^2, t = 2c - 1, cc = c (1 - c)
//   fmix = cc P,  fmix' = cc dP - t P,  fmix'' = cc (8 L_2 + 24 L_3 t) - 2 t dP - 2 P
// (the un-fused forms fmix / fmix_deriv / fmix_deriv2 above follow the reference's expressions term by term
//  and stay in use for the piecewise Strategy kernels; the per-cell Newton is FP64-pipe bound and evaluates
//  these 2-3 times per cell and evaluation: 9 instead of 26 FP64 instructions for fmix', 4 more instead of
//  25 for fmix'').  The results differ from the un-fused ones in the last bits; c_l, c_a agree with the
//  restatement to ~1e-15 either way, and what amplifies that (the composition RHS) is judged by the
//  extended-precision arbiter (tests/parity.py).
struct RKPieces {
   double t, cc, P, dP, d1;  // d1 = fmix'
};
AMPE_DEV RKPieces rk_eval(const double* L, double c)
{
   RKPieces r;
   r.t = fma(2.0, c, -1.0);
   r.cc = fma(-c, c, c);
   r.P = fma(r.t, fma(r.t, fma(r.t, L[3], L[2]), L[1]), L[0]);
   r.dP = fma(r.t, fma(r.t, 6.0 * L[3], 4.0 * L[2]), 2.0 * L[1]);
   r.d1 = fma(r.cc, r.dP, -(r.t * r.P));
   return r;
}
AMPE_DEV double rk_deriv2(const double* L, const RKPieces& r)
{
   const double d2P = fma(24.0 * L[3], r.t, 8.0 * L[2]);
   return fma(r.cc, d2P, fma(-2.0 * r.t, r.dP, -2.0 * r.P));
}

// what the Newton hands to the driving force: logarithms and polynomial pieces of the FINAL iterate
struct KksFinal {
   double lg[4];      // log(c_l), log(1-c_l), log(c_a), log(1-c_a) (unused where the argument is <= 1e-8)
   RKPieces rl, ra;
};

// KKS: (1-h) c_l + h c_a = c0,  mu_l(c_l) = mu_a(c_a)  (scaled by 1/RT), Cramer update,
// stop when both |F_i| < tol.  Returns iteration count, -1 if not converged.
AMPE_DEV int kks_newton(const CalphadT& t, double c0, double hphi, double& cl, double& ca,
                        double tol, int max_its, double alpha, KksFinal& F)
{
   c0 = c0 >= 0. ? c0 : 0.;
   c0 = c0 <= 1. ? c0 : 1.;
   const double dfab0 = t.fA[0] - t.fB[0], dfab1 = t.fA[1] - t.fB[1];
   int it = 0;
   while (true) {
      F.rl = rk_eval(t.L[0], cl);
      F.ra = rk_eval(t.L[1], ca);
      const double xi0 = t.RTinv * (dfab0 + F.rl.d1);
      const double xi1 = t.RTinv * (dfab1 + F.ra.d1);
      const double f0 = -c0 + (1.0 - hphi) * cl + hphi * ca;
      // xlogx_deriv(x) = log(x) + 1 above the 1e-8 extension.  Branch-free: the four logarithms are taken of a
      // normal argument in every lane (log_fast, fastmath.cuh: < 1 ulp, no range handling) and interleave;
      // the extension is a select.  AMPE_KKS_LIBM_LOG switches back to CUDA's log() (A/B builds).
      const double a0 = cl, a1 = 1. - cl, a2 = ca, a3 = 1. - ca;
      const bool b0 = a0 > AMPE_SMALLX, b1 = a1 > AMPE_SMALLX, b2 = a2 > AMPE_SMALLX, b3 = a3 > AMPE_SMALLX;
      F.lg[0] = AMPE_KKS_LOG(b0 ? a0 : 1.0);
      F.lg[1] = AMPE_KKS_LOG(b1 ? a1 : 1.0);
      F.lg[2] = AMPE_KKS_LOG(b2 ? a2 : 1.0);
      F.lg[3] = AMPE_KKS_LOG(b3 ? a3 : 1.0);
      const double d0 = b0 ? F.lg[0] + 1.0 : xlogx_deriv_ext(a0);
      const double d1 = b1 ? F.lg[1] + 1.0 : xlogx_deriv_ext(a1);
      const double d2 = b2 ? F.lg[2] + 1.0 : xlogx_deriv_ext(a2);
      const double d3 = b3 ? F.lg[3] + 1.0 : xlogx_deriv_ext(a3);
      const double f1 = d0 - d1 - d2 + d3 + (xi0 - xi1);
      if (fabs(f0) < tol && fabs(f1) < tol) return it;
      if (it == max_its) return -1;
      const double dxi0 = t.RTinv * rk_deriv2(t.L[0], F.rl);
      const double dxi1 = t.RTinv * rk_deriv2(t.L[1], F.ra);
      const double J00 = (1.0 - hphi), J01 = hphi;
      // the Jacobian only steers the iteration (the converged values are fixed by the residual): its
      // reciprocals are the straight-line ones
      const double J10 = dxi0 + xlogx_deriv2_fast(a0, b0) + xlogx_deriv2_fast(a1, b1);
      const double J11 = -dxi1 - xlogx_deriv2_fast(a2, b2) - xlogx_deriv2_fast(a3, b3);
      const double D = fma(J00, J11, -(J01 * J10));
      const double Dinv = AMPE_KKS_RCP(D);
      const double D0 = fma(f0, J11, -(J01 * f1));
      const double D1 = fma(J00, f1, -(f0 * J10));
      cl = fma(-alpha, Dinv * D0, cl);
      ca = fma(-alpha, Dinv * D1, ca);
      it++;
   }
}

// (f_l - f_a) - mu (c_l - c_a) with f_i = f(c_i) 1e-6/V_m, mu = df_a/dc(c_a) 1e-6/V_m
// (CALPHADFreeEnergyStrategyBinary.cc:321-323 computeFreeEnergy, 638-663 addDrivingForce): logarithms and
// Redlich-Kister pieces taken from the Newton's final residual, no further transcendental
AMPE_DEV double calphad_driving_force(const CalphadT& t, double cl, double ca, const KksFinal& F,
                                      double inv_vm_l, double inv_vm_a)
{
   const double* lg = F.lg;
   const double a0 = cl, a1 = 1.0 - cl, a2 = ca, a3 = 1.0 - ca;
   const double x0 = (a0 > AMPE_SMALLX) ? a0 * lg[0] : xlogx(a0, AMPE_LOG_SMALLX);
   const double x1 = (a1 > AMPE_SMALLX) ? a1 * lg[1] : xlogx(a1, AMPE_LOG_SMALLX);
   const double x2 = (a2 > AMPE_SMALLX) ? a2 * lg[2] : xlogx(a2, AMPE_LOG_SMALLX);
   const double x3 = (a3 > AMPE_SMALLX) ? a3 * lg[3] : xlogx(a3, AMPE_LOG_SMALLX);
   const double d2 = (a2 > AMPE_SMALLX) ? lg[2] + 1.0 : xlogx_deriv_ext(a2);
   const double d3 = (a3 > AMPE_SMALLX) ? lg[3] + 1.0 : xlogx_deriv_ext(a3);
   double f_l = fma(cl, t.fA[0], a1 * t.fB[0]) + fma(F.rl.cc, F.rl.P, t.RT * (x0 + x1));
   f_l *= inv_vm_l;
   double f_a = fma(ca, t.fA[1], a3 * t.fB[1]) + fma(F.ra.cc, F.ra.P, t.RT * (x2 + x3));
   f_a *= inv_vm_a;
   double mu = (t.fA[1] - t.fB[1]) + F.ra.d1 + t.RT * (d2 - d3);
   mu *= inv_vm_a;
   return (f_l - f_a) - mu * (cl - ca);
}

// computeDiffusionMobilityBinaryPhase (CALPHADMobility.cc:200-219) with
// getAtomicMobility / getDeltaG (CALPHADMobility.h:108-120, .cc:158-168)
AMPE_DEV double diffusion_mobility(const CalphadT& t, int phase, double c0)
{
   const double c1 = 1. - c0;
   const double dc = c0 - c1;
   double m[2];
#pragma unroll
   for (int sp = 0; sp < 2; sp++) {
      const double* qq = t.qAB[sp][phase];
      const double dG = c0 * t.qA[sp][phase] + c1 * t.qB[sp][phase] +
                        c0 * c1 * (qq[0] + dc * (qq[1] + dc * (qq[2] + dc * qq[3])));
      m[sp] = exp(dG * t.RTinv) * t.RTinv;
   }
   const double mm = c0 * m[1] + c1 * m[0];
   return c0 * c1 * mm * 1.e12;
}

}  // namespace ampe
