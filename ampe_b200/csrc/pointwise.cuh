// Pointwise device functions of the RHS path: interpolation / well / average
// polynomials (reference: source/fortran/functions.f:17-402), 1/|grad q| floor
// (source/fortran/quat.f:1497-1537) and the quaternion mobility
// (source/fortran/3d/mobility.m4:13-98).  Selector characters are uniform across
// a launch, so the switch is a uniform branch (the reference compares strings
// per cell).
#pragma once
#include <cuda_runtime.h>

namespace ampe {

#define AMPE_DEV __device__ __forceinline__

// max/min as compare + select (the state is never NaN; fmax/fmin expand to 5-7 instructions
// for their NaN rules, profiles/README.md)
AMPE_DEV double max0(double x) { return (x > 0.0) ? x : 0.0; }
AMPE_DEV double min1(double x) { return (x < 1.0) ? x : 1.0; }
AMPE_DEV double clamp01(double x) { return max0(min1(x)); }

// Rarely selected forms are kept out of line: inlining their log/cosh/tanh expansions at
// every call site bloated the fused kernel to ~12k SASS lines and showed up as
// instruction-fetch stalls (profiles/r01_dendrite2d.md).
#define AMPE_DEV_NOINLINE __device__ __noinline__
static AMPE_DEV_NOINLINE double interp_func_rare(double phi, char type)
{
   double phit;
   switch (type) {
      case 'w': phit = max0(phi); return phit * phit * (2.0 - phit);
      case 'm':
         phit = clamp01(phi);
         return phit * phit / (2.0 * phit * (phit - 1.0) + 1.0);
      case '3': phit = max0(phi); return phit * phit * phit;
      case 's': return log(cosh(10.0 * phi)) / log(cosh(10.0));  // unclamped: functions.f:74-77
      default: return 1.0;  // 'c'
   }
}
// interp_func, functions.f:17-91
AMPE_DEV double interp_func(double phi, char type)
{
   double phit;
   if (type == 'p') {
      phit = clamp01(phi);
      return phit * phit * phit * (10.0 - 15.0 * phit + 6.0 * phit * phit);
   }
   if (type == 'q') {
      phit = max0(phi);
      return phit * phit;
   }
   if (type == 'c') return 1.0;
   if (type == 'h') {
      phit = clamp01(phi);
      return phit * phit * (3.0 - 2.0 * phit);
   }
   if (type == 'l') {
      phit = max0(phi);
      return min1(phit);
   }
   return interp_func_rare(phi, type);
}

static AMPE_DEV_NOINLINE double deriv_interp_func_rare(double phi, char type)
{
   double phit, tmp;
   switch (type) {
      case 'w': phit = max0(phi); return phit * (4.0 - 3.0 * phit);
      case 'm':
         phit = clamp01(phi);
         tmp = 2.0 * phit * (phit - 1.0) + 1.0;
         return 2.0 * phit * (1.0 - phit) / (tmp * tmp);
      case '3': phit = max0(phi); return 3.0 * phit * phit;
      case 's': phit = max0(phi); return 10.0 * tanh(10.0 * phit) / log(cosh(10.0));
      default: return 0.0;  // 'c'
   }
}
// deriv_interp_func, functions.f:95-172
AMPE_DEV double deriv_interp_func(double phi, char type)
{
   double phit;
   if (type == 'p') {
      phit = clamp01(phi);
      return 30.0 * phit * phit * (1.0 - phit) * (1.0 - phit);
   }
   if (type == 'q') {
      phit = max0(phi);
      return 2.0 * phit;
   }
   if (type == 'c') return 0.0;
   if (type == 'h') {
      phit = clamp01(phi);
      return 6.0 * phit * (1.0 - phit);
   }
   if (type == 'l') return (phi > 0.0 || phi < 1.0) ? 1.0 : 0.0;  // `.or.`: functions.f:134-141
   return deriv_interp_func_rare(phi, type);
}

// deriv_well_func('d'|'s'), functions.f:276-300
AMPE_DEV double deriv_well_func(double phi, char type)
{
   if (type == 'd') return 32.0 * phi * (1.0 - phi) * (1.0 - 2.0 * phi);
   return 2.0 * (phi - 1.0);
}
AMPE_DEV double well_func(double phi, char type)
{
   if (type == 'd') return 16.0 * phi * phi * (1.0 - phi) * (1.0 - phi);
   return (1.0 - phi) * (1.0 - phi);
}

// average_func, functions.f:333-367
AMPE_DEV double average_func(double a, double b, char type)
{
   if (type == 'a') return 0.5 * (a + b);
   if (a < 1.0e-16 || b < 1.0e-16) return 0.0;
   return 2.0 / (1.0 / a + 1.0 / b);
}

static AMPE_DEV_NOINLINE double eval_grad_normi_rare(double g2, char floor_type, double floor2,
                                                     double max_normi)
{
   if (floor_type == 't') {
      const double gng2 = g2 * max_normi * max_normi;
      if (gng2 > 0.01) {
         const double gn = sqrt(g2);
         return tanh(max_normi * gn) / gn;
      }
      return max_normi * (1.0 - gng2 * (5.0 - 2.0 * gng2) / 15.0);
   }
   return 1.0 / sqrt(g2 + floor2);  // 's'
}
// eval_grad_normi, quat.f:1497-1537 (x**(-0.5) evaluated as 1/sqrt(x))
AMPE_DEV double eval_grad_normi(double g2, char floor_type, double floor2, double max_normi)
{
   if (floor_type == 'm') return (g2 > floor2) ? 1.0 / sqrt(g2) : max_normi;
   return eval_grad_normi_rare(g2, floor_type, floor2, max_normi);
}

static AMPE_DEV_NOINLINE double quat_mobility_rare(double phi, char func, double scale, double minm,
                                                   double alt)
{
   double qfunc;
   if (func == 'e' || func == 'E') {
      phi = clamp01(phi);
      qfunc = (1.0 - exp(alt * phi)) / (1.0 - exp(alt));
      qfunc = 1.0 - qfunc;
   } else {
      phi = fmax(1.e-6, fmin(1.0, phi));
      qfunc = fmax(0.0, (1.0 - phi) / (phi * phi));
      qfunc = fmin(qfunc, alt);
   }
   return minm + (scale - minm) * qfunc;
}
// quatmobility, 3d/mobility.m4:42-92
AMPE_DEV double quat_mobility(double phi, char func, double scale, double minm, double alt)
{
   if (func == 'p' || func == 'P') {
      phi = clamp01(phi);
      double qfunc = phi * phi * phi * (10.0 - 15.0 * phi + 6.0 * phi * phi);
      qfunc = 1.0 - qfunc;
      return minm + (scale - minm) * qfunc;
   }
   return quat_mobility_rare(phi, func, scale, minm, alt);
}

// quatmult4 / quatmult2, quat.f:867-911
AMPE_DEV void quatmult4(const double* a, const double* b, double* q)
{
   q[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
   q[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
   q[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
   q[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

}  // namespace ampe
