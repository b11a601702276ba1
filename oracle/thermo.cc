// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Stand-in for the part of ORNL Thermo4PFM that AMPE's RHS calls.  Thermo4PFM
// is an external dependency that is NOT in the reference tree and whose version
// is not pinned (README.md:28, CMakeLists.txt:280-290); this file restates its
// published algorithm:
//   * CALPHAD binary free energy  f = c G_A + (1-c) G_B + c(1-c) sum L_k (2c-1)^k
//     + RT [c ln c + (1-c) ln(1-c)]      (doc/latex/manual/appendix.tex:517-599)
//   * xlogx with a C2 quadratic extension below 1e-8
//   * KKS 2x2 Newton on  (1-h) c_l + h c_a = c,  mu_l(c_l) = mu_a(c_a), scaled by
//     1/RT, Cramer's rule update, |F_i| < tol stop test, tol=1e-8, alpha=1
//     (structure visible in-tree: source/CALPHADEqConcSolverBinaryWithPenalty.cc:25-126,
//      source/CALPHADConcSolverBinaryWithPenalty.cc:17-40)
//   * two-phase equilibrium (common tangent) Newton
//   * Quadratic free energies  f_i = A_i (c - ceq_i(T))^2 and the closed-form
//     KKS split (appendix.tex:426-490)
// Pinned by: tests/CALPHADbinaryEquilibrium/test.input (golden c_eq at 1423 K),
// tests/testCALPHADbinaryKKS.cc (mu_l == mu_a), tests/testCALPHADFunctions.cc
// (analytic vs FD).  Newton iterates themselves: parity unpinned.
#include "oracle.h"
#include <cmath>

namespace oracle {

// code comment CALPHADFreeEnergyStrategyBinary.cc:57 -- 8.314472 J/K/mol
const double GASCONSTANT_R_JPKPMOL = 8.314472;

static const double s_smallx = 1.0e-8;
static const double s_inv_smallx = 1. / s_smallx;
static const double s_log_smallx = log(s_smallx);
static const double s_smallx_log_smallx = s_smallx * s_log_smallx;
static const double s_one_plus_log_smallx = 1. + s_log_smallx;

double xlogx(double x)
{
   if (x > s_smallx) return x * log(x);
   return s_smallx_log_smallx + (x - s_smallx) * s_one_plus_log_smallx +
          0.5 * (x * x * s_inv_smallx - 2.0 * x + s_smallx);
}
double xlogx_deriv(double x)
{
   if (x > s_smallx) return log(x) + 1.0;
   return s_one_plus_log_smallx + (x - s_smallx) * s_inv_smallx;
}
double xlogx_deriv2(double x)
{
   if (x > s_smallx) return 1. / x;
   return s_inv_smallx;
}

double calphad_fmix(double l0, double l1, double l2, double l3, double c)
{
   const double t = 2.0 * c - 1.0;
   return c * (1.0 - c) * (l0 + l1 * t + l2 * t * t + l3 * t * t * t);
}
double calphad_fmix_deriv(double l0, double l1, double l2, double l3, double c)
{
   const double t = 2.0 * c - 1.0;
   const double cc = c * (1. - c);
   return (1.0 - 2.0 * c) * (l0 + l1 * t + l2 * t * t + l3 * t * t * t) +
          cc * (2.0 * l1 + 4.0 * l2 * t + 6.0 * l3 * t * t);
}
double calphad_fmix_deriv2(double l0, double l1, double l2, double l3, double c)
{
   const double t = 2.0 * c - 1.0;
   const double cc = c * (1. - c);
   return -2.0 * (l0 + l1 * t + l2 * t * t + l3 * t * t * t) +
          2.0 * (1.0 - 2.0 * c) * (2.0 * l1 + 4.0 * l2 * t + 6.0 * l3 * t * t) +
          cc * (8.0 * l2 + 24.0 * l3 * t);
}

// thermodynamic_data/calphadAuNi.dat:1-6
double calphad_species_fenergy(const ampe_calphad_species& s, double T)
{
   int iv = s.nintervals - 1;
   for (int i = 0; i < s.nintervals; i++)
      if (T >= s.Tc[i] && T < s.Tc[i + 1]) {
         iv = i;
         break;
      }
   if (T < s.Tc[0]) iv = 0;
   const double T2 = T * T;
   const double T4 = T2 * T2;
   return s.a[iv] + s.b[iv] * T + s.c[iv] * T * log(T) + s.d2[iv] * T2 + s.d3[iv] * T2 * T +
          s.d4[iv] * T4 + s.d7[iv] * T4 * T2 * T + s.dm1[iv] / T +
          s.dm9[iv] / (T4 * T4 * T);
}

void calphad_Tdep(const ampe_calphad_binary& db, double T, CalphadT& o)
{
   for (int ph = 0; ph < 2; ph++) {
      o.fA[ph] = calphad_species_fenergy(db.g[0][ph], T);
      o.fB[ph] = calphad_species_fenergy(db.g[1][ph], T);
      for (int k = 0; k < 4; k++) o.L[ph][k] = db.L[ph][k][0] + db.L[ph][k][1] * T;
   }
}

double calphad_free_energy(const ampe_calphad_binary& db, double T, double c, int pi)
{
   CalphadT t;
   calphad_Tdep(db, T, t);
   const double* L = t.L[pi];
   return c * t.fA[pi] + (1.0 - c) * t.fB[pi] + calphad_fmix(L[0], L[1], L[2], L[3], c) +
          GASCONSTANT_R_JPKPMOL * T * (xlogx(c) + xlogx(1.0 - c));
}
double calphad_deriv_free_energy(const ampe_calphad_binary& db, double T, double c, int pi)
{
   CalphadT t;
   calphad_Tdep(db, T, t);
   const double* L = t.L[pi];
   return (t.fA[pi] - t.fB[pi]) + calphad_fmix_deriv(L[0], L[1], L[2], L[3], c) +
          GASCONSTANT_R_JPKPMOL * T * (xlogx_deriv(c) - xlogx_deriv(1.0 - c));
}
double calphad_second_deriv_free_energy(const ampe_calphad_binary& db, double T, double c,
                                        int pi)
{
   CalphadT t;
   calphad_Tdep(db, T, t);
   const double* L = t.L[pi];
   return calphad_fmix_deriv2(L[0], L[1], L[2], L[3], c) +
          GASCONSTANT_R_JPKPMOL * T * (xlogx_deriv2(c) + xlogx_deriv2(1.0 - c));
}

// ---- 2x2 damped Newton with Cramer's rule ----------------------------------
template <class F>
static int newton2(F& sys, double* x, double tol, int max_its, double alpha)
{
   int it = 0;
   bool converged = false;
   while (true) {
      double fvec[2], J[2][2];
      sys.rhs(x, fvec);
      if (fabs(fvec[0]) < tol && fabs(fvec[1]) < tol) {
         converged = true;
         break;
      }
      if (it == max_its) break;
      sys.jac(x, J);
      const double D = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double Dinv = 1.0 / D;
      const double D0 = fvec[0] * J[1][1] - J[0][1] * fvec[1];
      const double D1 = J[0][0] * fvec[1] - fvec[0] * J[1][0];
      x[0] = x[0] - alpha * (Dinv * D0);
      x[1] = x[1] - alpha * (Dinv * D1);
      it++;
   }
   return converged ? it : -1;
}

struct KKSSystem {
   double c0, hphi, RTinv;
   CalphadT t;
   void xi(const double* c, double* out) const
   {
      for (int i = 0; i < 2; i++)
         out[i] = RTinv * (t.fA[i] - t.fB[i] +
                           calphad_fmix_deriv(t.L[i][0], t.L[i][1], t.L[i][2], t.L[i][3], c[i]));
   }
   void rhs(const double* c, double* f) const
   {
      double x[2];
      xi(c, x);
      f[0] = -c0 + (1.0 - hphi) * c[0] + hphi * c[1];
      f[1] = xlogx_deriv(c[0]) - xlogx_deriv(1. - c[0]) - xlogx_deriv(c[1]) +
             xlogx_deriv(1. - c[1]) + (x[0] - x[1]);
   }
   void jac(const double* c, double J[2][2]) const
   {
      double dxi[2];
      for (int i = 0; i < 2; i++)
         dxi[i] = RTinv *
                  calphad_fmix_deriv2(t.L[i][0], t.L[i][1], t.L[i][2], t.L[i][3], c[i]);
      J[0][0] = (1.0 - hphi);
      J[0][1] = hphi;
      J[1][0] = dxi[0] + xlogx_deriv2(c[0]) + xlogx_deriv2(1. - c[0]);
      J[1][1] = -dxi[1] - xlogx_deriv2(c[1]) - xlogx_deriv2(1. - c[1]);
   }
};

int calphad_phase_concentrations(const ampe_calphad_binary& db, double T, double c0,
                                 double hphi, double* x, double tol, int max_its,
                                 double alpha)
{
   KKSSystem s;
   s.RTinv = 1.0 / (GASCONSTANT_R_JPKPMOL * T);
   calphad_Tdep(db, T, s.t);
   // conc could be outside of [0,1] in a trial step
   s.c0 = c0 >= 0. ? c0 : 0.;
   s.c0 = s.c0 <= 1. ? s.c0 : 1.;
   s.hphi = hphi;
   return newton2(s, x, tol, max_its, alpha);
}

struct EqSystem {
   double RTinv, RT;
   CalphadT t;
   double f(int i, double c) const
   {
      return c * t.fA[i] + (1.0 - c) * t.fB[i] +
             calphad_fmix(t.L[i][0], t.L[i][1], t.L[i][2], t.L[i][3], c) +
             RT * (xlogx(c) + xlogx(1.0 - c));
   }
   double mu(int i, double c) const
   {
      return (t.fA[i] - t.fB[i]) +
             calphad_fmix_deriv(t.L[i][0], t.L[i][1], t.L[i][2], t.L[i][3], c) +
             RT * (xlogx_deriv(c) - xlogx_deriv(1.0 - c));
   }
   double d2(int i, double c) const
   {
      return calphad_fmix_deriv2(t.L[i][0], t.L[i][1], t.L[i][2], t.L[i][3], c) +
             RT * (xlogx_deriv2(c) + xlogx_deriv2(1.0 - c));
   }
   // common tangent, scaled by 1/RT
   void rhs(const double* c, double* fv) const
   {
      fv[0] = RTinv * (f(0, c[0]) - f(1, c[1]) - (c[0] - c[1]) * mu(1, c[1]));
      fv[1] = RTinv * (mu(0, c[0]) - mu(1, c[1]));
   }
   void jac(const double* c, double J[2][2]) const
   {
      J[0][0] = RTinv * (mu(0, c[0]) - mu(1, c[1]));
      J[0][1] = RTinv * (-(c[0] - c[1]) * d2(1, c[1]));
      J[1][0] = RTinv * d2(0, c[0]);
      J[1][1] = -RTinv * d2(1, c[1]);
   }
};

int calphad_ceq(const ampe_calphad_binary& db, double T, double* ceq, double tol,
                int max_its, double alpha)
{
   EqSystem s;
   s.RT = GASCONSTANT_R_JPKPMOL * T;
   s.RTinv = 1.0 / s.RT;
   calphad_Tdep(db, T, s.t);
   return newton2(s, ceq, tol, max_its, alpha);
}

// ---- quadratic ---------------------------------------------------------------
double quadratic_free_energy(const Quadratic& p, double T, double c, int pi)
{
   const double ceq = p.Ceq[pi] + (T - p.Tref) * p.m[pi];
   return p.A[pi] * (c - ceq) * (c - ceq);
}
double quadratic_deriv_free_energy(const Quadratic& p, double T, double c, int pi)
{
   const double ceq = p.Ceq[pi] + (T - p.Tref) * p.m[pi];
   return 2. * p.A[pi] * (c - ceq);
}
// appendix.tex:462-490 with h_eta = 0
void quadratic_phase_concentrations(const Quadratic& p, double T, double c0, double hphi,
                                    double* x)
{
   hphi = fmax(0.0, fmin(1.0, hphi));  // ConcInterpolationType::LINEAR inside Thermo4PFM
   const double ceql = p.Ceq[0] + (T - p.Tref) * p.m[0];
   const double ceqa = p.Ceq[1] + (T - p.Tref) * p.m[1];
   const double rla = p.A[0] / p.A[1];
   const double ral = p.A[1] / p.A[0];
   x[0] = (c0 - hphi * (ceqa - rla * ceql)) / ((1.0 - hphi) + hphi * rla);
   x[1] = (c0 - (1.0 - hphi) * (ceql - ral * ceqa)) / ((1.0 - hphi) * ral + hphi);
}

// ---- CALPHADMobility.{h,cc} (in-tree) ------------------------------------------
static double getQ(const double* a, double T)
{
   return a[0] + GASCONSTANT_R_JPKPMOL * T * log(a[1]);
}
// getDeltaG (CALPHADMobility.cc:158-168), getAtomicMobility (CALPHADMobility.h:108-120)
static double atomic_mobility(const ampe_calphad_binary& db, int species, int phase, double c0,
                              double c1, double T)
{
   const double dc = c0 - c1;
   const double qq0 = getQ(db.qAB[species][phase][0], T);
   const double qq1 = getQ(db.qAB[species][phase][1], T);
   const double qq2 = getQ(db.qAB[species][phase][2], T);
   const double qq3 = getQ(db.qAB[species][phase][3], T);
   const double dG = c0 * getQ(db.qA[species][phase], T) + c1 * getQ(db.qB[species][phase], T) +
                     c0 * c1 * (qq0 + dc * (qq1 + dc * (qq2 + dc * qq3)));
   const double rtinv = 1. / (GASCONSTANT_R_JPKPMOL * T);
   return exp(dG * rtinv) * rtinv;
}
// computeDiffusionMobilityBinaryPhase (CALPHADMobility.cc:200-219), m2toum2 = 1e12
double calphad_diffusion_mobility_binary(const ampe_calphad_binary& db, int phase, double c0,
                                         double T)
{
   const double c1 = 1. - c0;
   const double m0 = atomic_mobility(db, 0, phase, c0, c1, T);
   const double m1 = atomic_mobility(db, 1, phase, c0, c1, T);
   const double mm = c0 * m1 + c1 * m0;
   return c0 * c1 * mm * 1.e12;
}

}  // namespace oracle
