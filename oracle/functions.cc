// TEST INFRASTRUCTURE ONLY (see oracle.h).  Restatement of
// source/fortran/functions.f and the quaternion helpers of source/fortran/quat.f.
#include "oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace oracle {

static void die(const char* msg)
{
   fprintf(stderr, "%s\n", msg);
   abort();
}

// functions.f:17-91
double interp_func(double phi, char type)
{
   const double gamma = 10.0;
   double phit;
   switch (type) {
      case 'q':
         phit = fmax(0.0, phi);
         return phit * phit;
      case 'p':
         phit = fmax(0.0, fmin(1.0, phi));
         return phit * phit * phit * (10.0 - 15.0 * phit + 6.0 * phit * phit);
      case 'h':
         phit = fmax(0.0, fmin(1.0, phi));
         return phit * phit * (3.0 - 2.0 * phit);
      case 'w':
         phit = fmax(0.0, phi);
         return phit * phit * (2.0 - phit);
      case 'l':
         phit = fmax(0.0, phi);
         phit = fmin(1.0, phit);
         return phit;
      case 'm':
         phit = fmax(0.0, fmin(1.0, phi));
         return phit * phit / (2.0 * phit * (phit - 1.0) + 1.0);
      case '3':
         phit = fmax(0.0, phi);
         return phit * phit * phit;
      case 's':
         // functions.f:74-77: uses phi, not the clamped phit
         return log(cosh(gamma * phi)) / log(cosh(gamma));
      case 'c':
         return 1.0;
      default:
         die("Error in interp_func: type unknown");
   }
   return 0.0;
}

// functions.f:95-172
double deriv_interp_func(double phi, char type)
{
   const double gamma = 10.0;
   double phit, tmp;
   switch (type) {
      case 'q':
         phit = fmax(0.0, phi);
         return 2.0 * phit;
      case 'p':
         phit = fmax(0.0, fmin(1.0, phi));
         return 30.0 * phit * phit * (1.0 - phit) * (1.0 - phit);
      case 'h':
         phit = fmax(0.0, fmin(1.0, phi));
         return 6.0 * phit * (1.0 - phit);
      case 'w':
         phit = fmax(0.0, phi);
         return phit * (4.0 - 3.0 * phit);
      case 'l':
         // functions.f:134-141: `.or.` makes this 1 for every finite phi
         if (phi > 0.0 || phi < 1.0)
            return 1.0;
         else
            return 0.0;
      case 'm':
         phit = fmax(0.0, fmin(1.0, phi));
         tmp = 2.0 * phit * (phit - 1.0) + 1.0;
         return 2.0 * phit * (1.0 - phit) / (tmp * tmp);
      case '3':
         phit = fmax(0.0, phi);
         return 3.0 * phit * phit;
      case 's':
         phit = fmax(0.0, phi);
         return gamma * tanh(gamma * phit) / log(cosh(gamma));
      case 'c':
         return 0.0;
      default:
         die("Error in deriv_interp_func: unknown type");
   }
   return 0.0;
}

// functions.f:176-242
double second_deriv_interp_func(double phi, char type)
{
   double phit;
   switch (type) {
      case 'q':
         return (phi >= 0.0) ? 2.0 : 0.0;
      case 'p':
         phit = fmax(0.0, fmin(1.0, phi));
         return 60.0 * phit * (1.0 - 3.0 * phit + 2.0 * phit * phit);
      case 'h':
         phit = fmax(0.0, fmin(1.0, phi));
         return 6.0 * (1.0 - 2.0 * phit);
      case 'w':
         phit = fmax(0.0, phi);
         return 4.0 - 6.0 * phit;
      case 'l':
         return 0.0;
      case 'a':
         if (phi < 0.08333333333333333) {
            phit = fmax(0.0, phi);
            return 324.0 * phit;
         } else if (phi > 0.9166666666666666) {
            phit = fmin(1.0, phi);
            phit = 1.0 - phit;
            return -324.0 * phit;
         } else
            return 0.0;
      case 'c':
         return 0.0;
      default:
         die("Error in second_deriv_interp_func: type unknown");
   }
   return 0.0;
}

// functions.f:246-322
double well_func(double phi, char type)
{
   if (type == 'd') return 16.0 * phi * phi * (1.0 - phi) * (1.0 - phi);
   if (type == 's') return (1.0 - phi) * (1.0 - phi);
   die("Error in well_func: type unknown");
   return 0.0;
}
double deriv_well_func(double phi, char type)
{
   if (type == 'd') return 32.0 * phi * (1.0 - phi) * (1.0 - 2.0 * phi);
   if (type == 's') return 2.0 * (phi - 1.0);
   die("Error in deriv_well_func: type unknown");
   return 0.0;
}
double second_deriv_well_func(double phi, char type)
{
   if (type == 'd') return 32.0 * (1.0 + 6.0 * phi * (phi - 1.0));
   if (type == 's') return 2.0;
   die("Error in second_deriv_well_func: type unknown");
   return 0.0;
}

// functions.f:333-402
double average_func(double phi1, double phi2, char avg_type)
{
   const double threshold = 1.0e-16;
   if (avg_type == 'a') return 0.5 * (phi1 + phi2);
   if (avg_type == 'h') {
      if (phi1 < threshold || phi2 < threshold) return 0.0;
      return 2.0 / (1.0 / phi1 + 1.0 / phi2);
   }
   die("Error in average_func: type unknown");
   return 0.0;
}
double deriv_average_func(double avg_phi, double next_phi, char avg_type)
{
   const double threshold = 1.0e-16;
   if (avg_type == 'a') return 0.5;
   if (avg_type == 'h') {
      if (avg_phi < threshold) return 0.0;
      return 0.5 * next_phi * next_phi / (avg_phi * avg_phi);
   }
   die("Error in deriv_average_func: type unknown");
   return 0.0;
}

// functions.f: interp_ratio_func / compl_interp_ratio_func
double interp_ratio_func(double phi, char t1, char t2)
{
   if (t1 == t2) return 1.0;
   if (t1 == 'p' && t2 == 'l') {
      double phit = fmax(0.0, fmin(1.0, phi));
      return phit * phit * (10.0 - 15.0 * phit + 6.0 * phit * phit);
   }
   die("Error, interp_ratio: unknown/incompatible types");
   return 0.0;
}
double compl_interp_ratio_func(double phi, char t1, char t2)
{
   if (t1 == t2) return 1.0;
   if (t1 == 'p' && t2 == 'l') {
      double phit = fmax(0.0, fmin(1.0, phi));
      return (1.0 - phit) * (1.0 - phit) * (1.0 + 3.0 * phit + 6.0 * phit * phit);
   }
   die("compl_interp_ratio: unknown/incompatible types");
   return 0.0;
}

// ------------------------------------------------------------------ quat.f --
// quatmult4: quat.f:867-894
void quatmult4(const double* q1, const double* q2, double* q)
{
   q[0] = q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] - q1[3] * q2[3];
   q[1] = q1[0] * q2[1] + q1[1] * q2[0] + q1[2] * q2[3] - q1[3] * q2[2];
   q[2] = q1[0] * q2[2] + q1[2] * q2[0] + q1[3] * q2[1] - q1[1] * q2[3];
   q[3] = q1[0] * q2[3] + q1[3] * q2[0] + q1[1] * q2[2] - q1[2] * q2[1];
}
// quatmult2: quat.f:898-911
void quatmult2(const double* q1, const double* q2, double* q)
{
   q[0] = q1[0] * q2[0] - q1[1] * q2[1];
   q[1] = q1[0] * q2[1] + q1[1] * q2[0];
}
// quatconj: quat.f:915-927
void quatconj(const double* q1, double* q2)
{
   q2[0] = q1[0];
   q2[1] = -q1[1];
   q2[2] = -q1[2];
   q2[3] = -q1[3];
}

// setqr: quat.f:165-286 (entries normalised by quatset -> quatnorm4)
static double s_qr4[48][4];
static int s_conj4[48];
static bool s_qr4_ready = false;
static void setqr()
{
   static const int raw[48][4] = {
       {1, 0, 0, 0},    {0, 1, 0, 0},    {0, 0, 1, 0},    {0, 0, 0, 1},
       {-1, 0, 0, 0},   {0, -1, 0, 0},   {0, 0, -1, 0},   {0, 0, 0, -1},
       {1, 1, 0, 0},    {1, 0, 1, 0},    {1, 0, 0, 1},    {0, 1, 1, 0},
       {0, 1, 0, 1},    {0, 0, 1, 1},    {-1, 1, 0, 0},   {-1, 0, 1, 0},
       {-1, 0, 0, 1},   {0, -1, 1, 0},   {0, -1, 0, 1},   {0, 0, -1, 1},
       {1, -1, 0, 0},   {1, 0, -1, 0},   {1, 0, 0, -1},   {0, 1, -1, 0},
       {0, 1, 0, -1},   {0, 0, 1, -1},   {-1, -1, 0, 0},  {-1, 0, -1, 0},
       {-1, 0, 0, -1},  {0, -1, -1, 0},  {0, -1, 0, -1},  {0, 0, -1, -1},
       {1, 1, 1, 1},    {-1, 1, 1, 1},   {1, -1, 1, 1},   {1, 1, -1, 1},
       {1, 1, 1, -1},   {-1, -1, 1, 1},  {-1, 1, -1, 1},  {-1, 1, 1, -1},
       {1, -1, -1, 1},  {1, -1, 1, -1},  {1, 1, -1, -1},  {1, -1, -1, -1},
       {-1, 1, -1, -1}, {-1, -1, 1, -1}, {-1, -1, -1, 1}, {-1, -1, -1, -1}};
   static const int conj[48] = {1,  6,  7,  8,  5,  2,  3,  4,  21, 22, 23, 30,
                                31, 32, 27, 28, 29, 24, 25, 26, 9,  10, 11, 18,
                                19, 20, 15, 16, 17, 12, 13, 14, 44, 48, 43, 42,
                                41, 45, 46, 47, 37, 36, 35, 33, 38, 39, 40, 34};
   for (int n = 0; n < 48; n++) {
      double q[4] = {(double)raw[n][0], (double)raw[n][1], (double)raw[n][2],
                     (double)raw[n][3]};
      // quatnorm4 -> quatmaginv4 -> quatmagn4 (quat.f:945-1085)
      double m = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      double minv = (m < 1.e-15) ? 0.0 : 1.0 / m;
      for (int c = 0; c < 4; c++) s_qr4[n][c] = q[c] * minv;
      s_conj4[n] = conj[n];
   }
   s_qr4_ready = true;
}
const double* qr_table4()
{
   if (!s_qr4_ready) setqr();
   return &s_qr4[0][0];
}

// quatsymmrotate{4,2,1}: quat.f:41-71, 290-341, 449-520, 628-700
void quatsymmrotate(const double* q, int iq, double* q_prime, int qlen)
{
   if (qlen == 4) {
      if (!s_qr4_ready) setqr();
      if (iq == 0 || iq > 48 || iq < -48) die("Error in quatsymmrotate4");
      if (iq < 0) iq = s_conj4[-iq - 1];
      if (iq == 1) {
         for (int m = 0; m < 4; m++) q_prime[m] = q[m];
      } else {
         quatmult4(q, s_qr4[iq - 1], q_prime);
      }
   } else if (qlen == 2) {
      static const double qr2[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
      static const int conj2[4] = {1, 4, 3, 2};
      if (iq == 0 || iq > 4 || iq < -4) die("Error in quatsymmrotate2");
      if (iq < 0) iq = conj2[-iq - 1];
      if (iq == 1) {
         q_prime[0] = q[0];
         q_prime[1] = q[1];
      } else {
         quatmult2(q, qr2[iq - 1], q_prime);
      }
   } else if (qlen == 1) {
      const double PI = acos(-1.0);
      const double qr1[9] = {0.0,       0.5 * PI,  -0.5 * PI, PI,       -PI,
                             1.5 * PI, -1.5 * PI, 2.0 * PI,  -2.0 * PI};
      static const int conj1[9] = {1, 3, 2, 5, 4, 7, 6, 9, 8};
      if (iq == 0 || iq > 9 || iq < -9) die("Error in quatsymmrotate1");
      if (iq < 0) iq = conj1[-iq - 1];
      q_prime[0] = q[0] + qr1[iq - 1];
   } else {
      die("Error in quatsymmrotate, qlen");
   }
}

// eval_grad_normi: quat.f:1497-1537
double eval_grad_normi(double grad_norm2, char floor_type, double floor_grad_norm2,
                       double max_grad_normi)
{
   const double tol_taylor2 = 0.01;
   if (floor_type == 'm') {
      if (grad_norm2 > floor_grad_norm2) return pow(grad_norm2, -0.5);
      return max_grad_normi;
   } else if (floor_type == 't') {
      const double gng2 = grad_norm2 * max_grad_normi * max_grad_normi;
      if (gng2 > tol_taylor2) {
         const double grad_norm = sqrt(grad_norm2);
         return tanh(max_grad_normi * grad_norm) / grad_norm;
      }
      return max_grad_normi * (1.0 - gng2 * (5.0 - 2.0 * gng2) / 15.0);
   } else if (floor_type == 's') {
      return pow(grad_norm2 + floor_grad_norm2, -0.5);
   }
   die("Error in eval_grad_normi: floor_type unknown");
   return 0.0;
}

}  // namespace oracle
