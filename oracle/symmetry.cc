// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the quaternion symmetry pre-pass and of the CVODE projection hook
// (SURVEY.md 8f ranks 1 and 4), routine by routine:
//   quatfindsymm{4,2,1}       quat.f:9-37, 73-163, 343-445, 524-624
//   quatdiffsq / rotatediffsq quat.f:1171-1310 (operands normalised with quatmaginv, :985-1039)
//   quat_symm_rotation        {2d,3d}/quatrotation.m4:12-89   (QuatModel.cc:4978-5055)
//   quat_fundamental          {2d,3d}/quatrotation.m4:93-147  (QuatModel.cc:5059-5104)
//   project{2,3}d             {2d,3d}/quatfacops.m4 project2d / 3d:1022-1081 (QuatFACOps.cc:2395-2434,
//                             QuatIntegrator::applyProjection, QuatIntegrator.cc:3911-3962)
// Parity status: the reference holds no known-answer test for these routines; the restatement is
// checked by brute-force properties in tests/test_oracle_symmetry.py ("parity unpinned" otherwise).
#include <cmath>
#include <cstdlib>

#include "oracle.h"

namespace oracle {

namespace {

// quatnorm4 / quatnorm2 on a copy (quatmaginv: 0 below 1e-15)
inline void norm_copy(const double* q, double* o, int n)
{
   double s = 0.0;
   for (int m = 0; m < n; m++) s = s + q[m] * q[m];
   const double mag = sqrt(s);
   const double minv = (mag < 1.e-15) ? 0.0 : 1.0 / mag;
   for (int m = 0; m < n; m++) o[m] = q[m] * minv;
}
// quatdiffsq{4,2}: |norm(q2) - norm(q1)|^2
inline double diffsq(const double* q1, const double* q2, int n)
{
   double a[4], b[4];
   norm_copy(q1, a, n);
   norm_copy(q2, b, n);
   double s = 0.0;
   for (int m = 0; m < n; m++) {
      const double d = b[m] - a[m];
      s = s + d * d;
   }
   return s;
}

const double qr2[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
const int conj2[4] = {1, 4, 3, 2};
const int conj1[9] = {1, 3, 2, 5, 4, 7, 6, 9, 8};
const int conj4[48] = {1,  6,  7,  8,  5,  2,  3,  4,  21, 22, 23, 30, 31, 32, 27, 28,
                       29, 24, 25, 26, 9,  10, 11, 18, 19, 20, 15, 16, 17, 12, 13, 14,
                       44, 48, 43, 42, 41, 45, 46, 47, 37, 36, 35, 33, 38, 39, 40, 34};

// candidate nn (1-based): q2 rotated, and its squared distance to q1
inline double candidate(const double* q1, const double* q2, int nn, double* q2p, int qlen)
{
   if (nn == 1) {
      for (int m = 0; m < qlen; m++) q2p[m] = q2[m];
   } else if (qlen == 4) {
      quatmult4(q2, qr_table4() + 4 * (nn - 1), q2p);
   } else {
      quatmult2(q2, qr2[nn - 1], q2p);
   }
   return diffsq(q1, q2p, qlen);
}

}  // namespace

// quatfindsymm (quat.f:9-37): iq is in/out -- the rotation found last time is tried first
void quatfindsymm(const double* q1, const double* q2, int* iq_io, double* q2_prime, int qlen)
{
   int iq = *iq_io;
   if (qlen == 4 || qlen == 2) {
      const int NROT = (qlen == 4) ? 48 : 4;
      const double PI = acos(-1.0);
      const double s = 2.0 * sin(PI / (qlen == 4 ? 16.0 : 8.0));
      const double thr = s * s;
      if (iq == 0 || iq > NROT || iq < -NROT) iq = 1;
      if (iq < 0) iq = (qlen == 4) ? conj4[-iq - 1] : conj2[-iq - 1];
      double dsq = candidate(q1, q2, iq, q2_prime, qlen);
      *iq_io = iq;
      if (dsq <= thr) return;
      double min_dsq = dsq;
      int min_iq = iq;
      double min_q2p[4];
      for (int m = 0; m < qlen; m++) min_q2p[m] = q2_prime[m];
      for (int nn = 1; nn <= NROT; nn++) {
         if (nn == iq) continue;
         dsq = candidate(q1, q2, nn, q2_prime, qlen);
         if (dsq < min_dsq) {
            min_dsq = dsq;
            min_iq = nn;
            for (int m = 0; m < qlen; m++) min_q2p[m] = q2_prime[m];
         }
         if (dsq <= thr) break;
      }
      for (int m = 0; m < qlen; m++) q2_prime[m] = min_q2p[m];
      *iq_io = min_iq;
   } else if (qlen == 1) {
      const double PI = acos(-1.0);
      const double qr1[9] = {0.0,      0.5 * PI,  -0.5 * PI, PI,       -PI,
                             1.5 * PI, -1.5 * PI, 2.0 * PI,  -2.0 * PI};
      const double PI_OVER_4 = 0.25 * PI;
      if (iq == 0 || iq > 9 || iq < -9) iq = 1;
      if (iq < 0) iq = conj1[-iq - 1];
      auto cand = [&](int nn, double* q2p) {
         *q2p = (nn == 1) ? q2[0] : q2[0] + qr1[nn - 1];
         return fabs(*q2p - q1[0]);
      };
      double d = cand(iq, q2_prime);
      *iq_io = iq;
      if (d <= PI_OVER_4) return;
      double min_d = d, min_q = *q2_prime;
      int min_iq = iq;
      for (int nn = 1; nn <= 9; nn++) {
         if (nn == iq) continue;
         d = cand(nn, q2_prime);
         if (d < min_d) {
            min_d = d;
            min_iq = nn;
            min_q = *q2_prime;
         }
         if (d <= PI_OVER_4) break;
      }
      *q2_prime = min_q;
      *iq_io = min_iq;
   } else {
      abort();
   }
}

// quat_symm_rotation: rot[a](face) <- rotation that brings the lower neighbour closest to the cell;
// x faces over j,k grown by one ghost, etc. (quatrotation.m4:37-86)
void quat_symm_rotation(const Box& b, View q, int depth, IView* rot)
{
   const int nd = b.ndim;
   for (int a = 0; a < nd; a++) {
      int L[3], H[3];
      for (int d = 0; d < 3; d++) {
         const int g = (d < nd && d != a) ? 1 : 0;
         L[d] = b.lo[d] - g;
         H[d] = b.hi[d] + g + (d == a ? 1 : 0);
      }
      for (int k = L[2]; k <= H[2]; k++)
         for (int j = L[1]; j <= H[1]; j++)
            for (int i = L[0]; i <= H[0]; i++) {
               double q1[4], q2[4], q2p[4];
               for (int m = 0; m < depth; m++) {
                  q1[m] = q(i, j, k, m);
                  q2[m] = q(i - (a == 0), j - (a == 1), k - (a == 2), m);
               }
               quatfindsymm(q1, q2, &rot[a](i, j, k), q2p, depth);
            }
   }
}

// quat_fundamental: every quaternion replaced by its symmetric equivalent closest to the identity
void quat_fundamental(const Box& b, View quat, int depth)
{
   double q1[4] = {0.0, 0.0, 0.0, 0.0};
   if (depth == 2 || depth == 4) q1[0] = 1.0;  // quatset -> normalised (1,0[,0,0])
   for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
         for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            double q2[4], q2p[4];
            for (int m = 0; m < depth; m++) q2[m] = quat(i, j, k, m);
            int iq = 1;
            quatfindsymm(q1, q2, &iq, q2p, depth);
            for (int m = 0; m < depth; m++) quat(i, j, k, m) = q2p[m];
         }
}

// project{2,3}d: corr <- q/|q| - q ; err <- err - (err . q/|q|) q/|q|
void project(const Box& b, int depth, View q, View corr, View err)
{
   for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
         for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            double fac = 0.0;
            for (int m = 0; m < depth; m++) fac = fac + q(i, j, k, m) * q(i, j, k, m);
            fac = 1.0 / sqrt(fac);
            for (int m = 0; m < depth; m++) corr(i, j, k, m) = q(i, j, k, m) * fac;
            fac = 0.0;
            for (int m = 0; m < depth; m++) fac = fac + corr(i, j, k, m) * err(i, j, k, m);
            for (int m = 0; m < depth; m++) err(i, j, k, m) = err(i, j, k, m) - corr(i, j, k, m) * fac;
            for (int m = 0; m < depth; m++) corr(i, j, k, m) = corr(i, j, k, m) - q(i, j, k, m);
         }
}

}  // namespace oracle
