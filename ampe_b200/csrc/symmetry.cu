// Quaternion symmetry pre-pass and the CVODE projection hook on the device (SURVEY.md 8f ranks 1
// and 4), behind the C ABI:
//   ampe_k_quat_symm_rotation   QUAT_SYMM_ROTATION  {2d,3d}/quatrotation.m4:12-89
//   ampe_k_quat_fundamental     QUAT_FUNDAMENTAL    {2d,3d}/quatrotation.m4:93-147
//   ampe_k_project              PROJECT{2,3}D       3d/quatfacops.m4:1022-1081
//   ampe_rhs_compute_symmetry_rotations   QuatModel::computeSymmetryRotations (QuatModel.cc:4978-5055)
//   ampe_quat_fundamental                 QuatModel::makeQuatFundamental      (QuatModel.cc:5059-5104)
//   ampe_apply_projection                 QuatIntegrator::applyProjection     (QuatIntegrator.cc:3911-3962)
// The search quatfindsymm{4,2,1} (quat.f:73-163, 343-445, 524-624) is restated with the reference's
// candidate order, early exit and operation order (the library is built with --fmad=false), so the
// integer rotation indices reproduce the CPU restatement of the reference bit for bit.  The 48 cubic rotations live in shared
// memory for the lifetime of a block.  One-off, bandwidth-trivial kernels (one thread per face / cell).
#include <cuda_runtime.h>

#include <string>

#include "../../include/ampe_b200_kernels.h"
#include "ctx_internal.h"
#include "pointwise.cuh"

namespace {
using namespace ampe;

#include "box_view.cuh"

typedef V<int> IWV;

#define CUDA_OKS(call)                                                                        \
   do {                                                                                      \
      cudaError_t e_ = (call);                                                               \
      if (e_ != cudaSuccess)                                                                 \
         return ampe_set_err(AMPE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
   } while (0)

// rotation table of setqr (quat.f:165-286) and the conjugate indices, per device
__constant__ double c_qr4[48 * 4];
__constant__ int c_conj4[48];
bool g_table_ready[64] = {false};

int ensure_table()
{
   int dev = 0;
   CUDA_OKS(cudaGetDevice(&dev));
   if (dev >= 0 && dev < 64 && g_table_ready[dev]) return AMPE_OK;
   static const int raw[48][4] = {
       {1, 0, 0, 0},    {0, 1, 0, 0},    {0, 0, 1, 0},    {0, 0, 0, 1},    {-1, 0, 0, 0},
       {0, -1, 0, 0},   {0, 0, -1, 0},   {0, 0, 0, -1},   {1, 1, 0, 0},    {1, 0, 1, 0},
       {1, 0, 0, 1},    {0, 1, 1, 0},    {0, 1, 0, 1},    {0, 0, 1, 1},    {-1, 1, 0, 0},
       {-1, 0, 1, 0},   {-1, 0, 0, 1},   {0, -1, 1, 0},   {0, -1, 0, 1},   {0, 0, -1, 1},
       {1, -1, 0, 0},   {1, 0, -1, 0},   {1, 0, 0, -1},   {0, 1, -1, 0},   {0, 1, 0, -1},
       {0, 0, 1, -1},   {-1, -1, 0, 0},  {-1, 0, -1, 0},  {-1, 0, 0, -1},  {0, -1, -1, 0},
       {0, -1, 0, -1},  {0, 0, -1, -1},  {1, 1, 1, 1},    {-1, 1, 1, 1},   {1, -1, 1, 1},
       {1, 1, -1, 1},   {1, 1, 1, -1},   {-1, -1, 1, 1},  {-1, 1, -1, 1},  {-1, 1, 1, -1},
       {1, -1, -1, 1},  {1, -1, 1, -1},  {1, 1, -1, -1},  {1, -1, -1, -1}, {-1, 1, -1, -1},
       {-1, -1, 1, -1}, {-1, -1, -1, 1}, {-1, -1, -1, -1}};
   static const int conj[48] = {1,  6,  7,  8,  5,  2,  3,  4,  21, 22, 23, 30, 31, 32, 27, 28,
                                29, 24, 25, 26, 9,  10, 11, 18, 19, 20, 15, 16, 17, 12, 13, 14,
                                44, 48, 43, 42, 41, 45, 46, 47, 37, 36, 35, 33, 38, 39, 40, 34};
   double qr[48 * 4];
   for (int n = 0; n < 48; n++) {
      // quatset -> quatnorm4 -> quatmaginv4 (quat.f:704-716, 985-1010, 1069-1083)
      const double q[4] = {(double)raw[n][0], (double)raw[n][1], (double)raw[n][2], (double)raw[n][3]};
      const double m = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      const double minv = (m < 1.e-15) ? 0.0 : 1.0 / m;
      for (int k = 0; k < 4; k++) qr[4 * n + k] = q[k] * minv;
   }
   CUDA_OKS(cudaMemcpyToSymbol(c_qr4, qr, sizeof(qr)));
   CUDA_OKS(cudaMemcpyToSymbol(c_conj4, conj, sizeof(conj)));
   if (dev >= 0 && dev < 64) g_table_ready[dev] = true;
   return AMPE_OK;
}

// thresholds of the search, formed like the reference does at its first pass
struct SymmConst {
   double thr4;      // (2 sin(pi/16))^2   quat.f:112-115
   double thr2;      // (2 sin(pi/8))^2    quat.f:383-386
   double pi;        // dacos(-1)
};
SymmConst symm_const()
{
   SymmConst s;
   s.pi = acos(-1.0);
   const double a = 2.0 * sin(s.pi / 16.0), b = 2.0 * sin(s.pi / 8.0);
   s.thr4 = a * a;
   s.thr2 = b * b;
   return s;
}

// quatnorm{4,2} of a copy: q * (1/|q|), 0 below 1e-15 (quatmaginv, quat.f:985-1039)
template <int Q>
__device__ __forceinline__ void norm_copy(const double* q, double* o)
{
   double s = 0.0;
#pragma unroll
   for (int m = 0; m < Q; m++) s = s + q[m] * q[m];
   const double mag = sqrt(s);
   const double minv = (mag < 1.e-15) ? 0.0 : 1.0 / mag;
#pragma unroll
   for (int m = 0; m < Q; m++) o[m] = q[m] * minv;
}

// candidate nn (1-based): q2p = q2 * qr(nn) (copy for nn = 1); returns |norm(q2p) - q1n|^2
// (quatdiffsq / quatrotatediffsq, quat.f:1171-1295; q1n = norm(q1) is the same for every candidate)
template <int Q>
__device__ __forceinline__ double candidate(const double* s_qr, const double* q1n, const double* q2, int nn,
                                            double* q2p)
{
   if (nn == 1) {
#pragma unroll
      for (int m = 0; m < Q; m++) q2p[m] = q2[m];
   } else if (Q == 4) {
      quatmult4(q2, s_qr + 4 * (nn - 1), q2p);
   } else {
      // quatmult2 with (1,0),(0,1),(-1,0),(0,-1)  (quat.f:389-392, 898-911)
      const double r0 = (nn == 3) ? -1.0 : ((nn == 1) ? 1.0 : 0.0);
      const double r1 = (nn == 2) ? 1.0 : ((nn == 4) ? -1.0 : 0.0);
      const double a = q2[0] * r0 - q2[1] * r1;
      const double b = q2[0] * r1 + q2[1] * r0;
      q2p[0] = a;
      q2p[1] = b;
   }
   double t[Q];
   norm_copy<Q>(q2p, t);
   double s = 0.0;
#pragma unroll
   for (int m = 0; m < Q; m++) {
      const double d = t[m] - q1n[m];
      s = s + d * d;
   }
   return s;
}

// quatfindsymm{4,2}: returns the rotation index; q2p = rotated q2
template <int Q>
__device__ int findsymm(const double* s_qr, const int* s_conj, const double* q1, const double* q2, int iq,
                        double* q2p, double thr)
{
   constexpr int NROT = (Q == 4) ? 48 : 4;
   if (iq == 0 || iq > NROT || iq < -NROT) iq = 1;
   if (iq < 0) iq = (Q == 4) ? s_conj[-iq - 1] : ((iq == -2) ? 4 : ((iq == -4) ? 2 : -iq));
   double q1n[Q];
   norm_copy<Q>(q1, q1n);
   double dsq = candidate<Q>(s_qr, q1n, q2, iq, q2p);
   if (dsq <= thr) return iq;
   double min_dsq = dsq;
   int min_iq = iq;
   double best[Q], tmp[Q];
#pragma unroll
   for (int m = 0; m < Q; m++) best[m] = q2p[m];
   for (int nn = 1; nn <= NROT; nn++) {
      if (nn == iq) continue;
      dsq = candidate<Q>(s_qr, q1n, q2, nn, tmp);
      if (dsq < min_dsq) {
         min_dsq = dsq;
         min_iq = nn;
#pragma unroll
         for (int m = 0; m < Q; m++) best[m] = tmp[m];
      }
      if (dsq <= thr) break;
   }
#pragma unroll
   for (int m = 0; m < Q; m++) q2p[m] = best[m];
   return min_iq;
}

// quatfindsymm1 (quat.f:524-624): the orientation is one angle, rotations are multiples of pi/2
__device__ int findsymm1(double q1, double q2, int iq, double* q2p, double pi)
{
   const double qr[9] = {0.0, 0.5 * pi, -(0.5 * pi), pi, -pi, 1.5 * pi, -(1.5 * pi), 2.0 * pi, -(2.0 * pi)};
   const int conj[9] = {1, 3, 2, 5, 4, 7, 6, 9, 8};
   const double quarter = 0.25 * pi;
   if (iq == 0 || iq > 9 || iq < -9) iq = 1;
   if (iq < 0) iq = conj[-iq - 1];
   double cur = (iq == 1) ? q2 : q2 + qr[iq - 1];
   double d = fabs(cur - q1);
   *q2p = cur;
   if (d <= quarter) return iq;
   double min_d = d, best = cur;
   int min_iq = iq;
   for (int nn = 1; nn <= 9; nn++) {
      if (nn == iq) continue;
      cur = (nn == 1) ? q2 : q2 + qr[nn - 1];
      d = fabs(cur - q1);
      if (d < min_d) {
         min_d = d;
         min_iq = nn;
         best = cur;
      }
      if (d <= quarter) break;
   }
   *q2p = best;
   return min_iq;
}

template <int Q>
__device__ __forceinline__ int findsymm_any(const double* s_qr, const int* s_conj, const double* q1,
                                            const double* q2, int iq, double* q2p, const SymmConst& sc)
{
   if constexpr (Q == 1)
      return findsymm1(q1[0], q2[0], iq, q2p, sc.pi);
   else
      return findsymm<Q>(s_qr, s_conj, q1, q2, iq, q2p, Q == 4 ? sc.thr4 : sc.thr2);
}

__device__ __forceinline__ void load_table(double* s_qr, int* s_conj)
{
   for (int t = threadIdx.x; t < 48 * 4; t += blockDim.x) s_qr[t] = c_qr4[t];
   for (int t = threadIdx.x; t < 48; t += blockDim.x) s_conj[t] = c_conj4[t];
   __syncthreads();
}

// ---- SAMRAI-layout kernels (piecewise boundary) ----------------------------------------------
// faces of axis A over the box [L,H]: rot(face) <- findsymm(q(cell), q(lower neighbour), rot(face))
template <int Q>
__global__ void symm_rotation_box_kernel(int L0, int L1, int L2, int e0, int e1, long long total, int a,
                                         CV q, IWV rot, SymmConst sc)
{
   __shared__ double s_qr[48 * 4];
   __shared__ int s_conj[48];
   load_table(s_qr, s_conj);
   for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
        t += (long long)gridDim.x * blockDim.x) {
      const int i = L0 + (int)(t % e0);
      const int j = L1 + (int)((t / e0) % e1);
      const int k = L2 + (int)(t / ((long long)e0 * e1));
      double q1[Q], q2[Q], q2p[Q];
#pragma unroll
      for (int m = 0; m < Q; m++) {
         q1[m] = q(i, j, k, m);
         q2[m] = q(i - E(a, 0), j - E(a, 1), k - E(a, 2), m);
      }
      rot(i, j, k) = findsymm_any<Q>(s_qr, s_conj, q1, q2, rot(i, j, k), q2p, sc);
   }
}

template <int Q>
__global__ void fundamental_box_kernel(int L0, int L1, int L2, int e0, int e1, long long total, DV quat,
                                       SymmConst sc)
{
   __shared__ double s_qr[48 * 4];
   __shared__ int s_conj[48];
   load_table(s_qr, s_conj);
   for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
        t += (long long)gridDim.x * blockDim.x) {
      const int i = L0 + (int)(t % e0);
      const int j = L1 + (int)((t / e0) % e1);
      const int k = L2 + (int)(t / ((long long)e0 * e1));
      double q1[Q], q2[Q], q2p[Q];
#pragma unroll
      for (int m = 0; m < Q; m++) {
         q1[m] = (m == 0 && Q > 1) ? 1.0 : 0.0;
         q2[m] = quat(i, j, k, m);
      }
      (void)findsymm_any<Q>(s_qr, s_conj, q1, q2, 1, q2p, sc);
#pragma unroll
      for (int m = 0; m < Q; m++) quat(i, j, k, m) = q2p[m];
   }
}

// project{2,3}d: corr <- q/|q| - q, err <- err - (err . q/|q|) q/|q|
template <int Q>
__global__ void project_box_kernel(int L0, int L1, int L2, int e0, int e1, long long total, CV q, DV corr,
                                   DV err)
{
   for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
        t += (long long)gridDim.x * blockDim.x) {
      const int i = L0 + (int)(t % e0);
      const int j = L1 + (int)((t / e0) % e1);
      const int k = L2 + (int)(t / ((long long)e0 * e1));
      double v[Q], c[Q], e[Q];
      double fac = 0.0;
#pragma unroll
      for (int m = 0; m < Q; m++) {
         v[m] = q(i, j, k, m);
         fac = fac + v[m] * v[m];
      }
      fac = 1.0 / sqrt(fac);
#pragma unroll
      for (int m = 0; m < Q; m++) c[m] = v[m] * fac;
      fac = 0.0;
#pragma unroll
      for (int m = 0; m < Q; m++) {
         e[m] = err(i, j, k, m);
         fac = fac + c[m] * e[m];
      }
#pragma unroll
      for (int m = 0; m < Q; m++) {
         err(i, j, k, m) = e[m] - c[m] * fac;
         corr(i, j, k, m) = c[m] - v[m];
      }
   }
}

int blocks_for(long long total)
{
   const long long b = (total + 255) / 256;
   return (int)(b > 148LL * 16 ? 148LL * 16 : (b < 1 ? 1 : b));
}

// view over an explicit ghost box [glo, ghi] (the reference passes lo/hi of the array instead of a
// ghost width for QUAT_FUNDAMENTAL and PROJECT)
template <typename T>
V<T> view_lohi(T* p, int ndim, const int* glo, const int* ghi)
{
   V<T> v;
   v.p = p;
   int n[3], lo[3];
   for (int d = 0; d < 3; d++) {
      lo[d] = d < ndim ? glo[d] : 0;
      n[d] = d < ndim ? ghi[d] - glo[d] + 1 : 1;
   }
   v.lo0 = lo[0], v.lo1 = lo[1], v.lo2 = lo[2];
   v.n0 = n[0], v.n1 = n[1];
   v.comp = (long long)n[0] * n[1] * n[2];
   return v;
}

int check_last(const char* what)
{
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return ampe_set_err(AMPE_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
   return AMPE_OK;
}

int launch_fundamental(int depth, const int* L, const int* H, DV quat, cudaStream_t st)
{
   const int e0 = H[0] - L[0] + 1, e1 = H[1] - L[1] + 1, e2 = H[2] - L[2] + 1;
   if (e0 <= 0 || e1 <= 0 || e2 <= 0) return AMPE_OK;
   int rc = ensure_table();
   if (rc) return rc;
   const long long total = (long long)e0 * e1 * e2;
   const SymmConst sc = symm_const();
   const int nb = blocks_for(total);
   if (depth == 4)
      fundamental_box_kernel<4><<<nb, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, quat, sc);
   else if (depth == 2)
      fundamental_box_kernel<2><<<nb, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, quat, sc);
   else if (depth == 1)
      fundamental_box_kernel<1><<<nb, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, quat, sc);
   else
      return ampe_set_err(AMPE_EINVAL, "quat_fundamental: depth must be 1, 2 or 4 (quatfindsymm stops)");
   return check_last("quat_fundamental");
}

int launch_symm_rotation(int depth, int a, const int* L, const int* H, CV q, IWV rot, cudaStream_t st)
{
   const int e0 = H[0] - L[0] + 1, e1 = H[1] - L[1] + 1, e2 = H[2] - L[2] + 1;
   if (e0 <= 0 || e1 <= 0 || e2 <= 0) return AMPE_OK;
   int rc = ensure_table();
   if (rc) return rc;
   const long long total = (long long)e0 * e1 * e2;
   const SymmConst sc = symm_const();
   const int nb = blocks_for(total);
   if (depth == 4)
      symm_rotation_box_kernel<4><<<nb, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, a, q, rot, sc);
   else if (depth == 2)
      symm_rotation_box_kernel<2><<<nb, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, a, q, rot, sc);
   else if (depth == 1)
      symm_rotation_box_kernel<1><<<nb, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, a, q, rot, sc);
   else
      return ampe_set_err(AMPE_EINVAL, "quat_symm_rotation: depth must be 1, 2 or 4 (quatfindsymm stops)");
   return check_last("quat_symm_rotation");
}

// ---- context-level kernels: ghost-0 arrays of the solution vector, periodic level ------------
// rot_a(cell) for the LOWER face of every interior cell; neighbours wrap periodically (single rank)
template <int Q, int ND>
__global__ void symm_rotation_ctx_kernel(int n0, int n1, int n2, const double* __restrict__ q, int* rot0,
                                         int* rot1, int* rot2, SymmConst sc, const double* __restrict__ q_lo,
                                         long long lo_comp, int ng)
{
   // q_lo != null (slab rank with neighbours): the ng planes below plane 0 along the slab axis (component
   // stride lo_comp) instead of the periodic wrap inside this rank
   __shared__ double s_qr[48 * 4];
   __shared__ int s_conj[48];
   load_table(s_qr, s_conj);
   const long long ncell = (long long)n0 * n1 * n2;
   for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ncell;
        t += (long long)gridDim.x * blockDim.x) {
      const int i = (int)(t % n0);
      const int j = (int)((t / n0) % n1);
      const int k = (int)(t / ((long long)n0 * n1));
      double q1[Q], q2[Q], q2p[Q];
#pragma unroll
      for (int m = 0; m < Q; m++) q1[m] = q[t + m * ncell];
#pragma unroll
      for (int a = 0; a < ND; a++) {
         const int ii = (a == 0) ? (i == 0 ? n0 - 1 : i - 1) : i;
         const int jj = (a == 1) ? (j == 0 ? n1 - 1 : j - 1) : j;
         const int kk = (a == 2) ? (k == 0 ? n2 - 1 : k - 1) : k;
         const long long nb = ii + (long long)n0 * (jj + (long long)n1 * kk);
         const bool below = q_lo && a == ND - 1 && ((ND == 3) ? k == 0 : j == 0);
         if (below) {
            // the neighbour's highest plane: the last of its ng ghost planes held here
            const long long pl = (ND == 3) ? (long long)n0 * n1 : n0;
            const long long o = (long long)(ng - 1) * pl + ((ND == 3) ? i + (long long)n0 * j : i);
#pragma unroll
            for (int m = 0; m < Q; m++) q2[m] = q_lo[o + m * lo_comp];
         } else {
#pragma unroll
            for (int m = 0; m < Q; m++) q2[m] = q[nb + m * ncell];
         }
         int* rot = (a == 0) ? rot0 : ((a == 1) ? rot1 : rot2);
         rot[t] = findsymm_any<Q>(s_qr, s_conj, q1, q2, rot[t], q2p, sc);
      }
   }
}

}  // namespace

extern "C" {

/* QUAT_SYMM_ROTATION (QuatFort.h:319; {2d,3d}/quatrotation.m4:12-89).  rot[a]: SideData<int> of axis
 * a with ghost width ngrot, IN/OUT (the previous index is tried first; 0 or out of range means 1).
 * Faces of axis a over the box grown by one cell in the transverse directions (needs ngq, ngrot >= 1). */
int ampe_k_quat_symm_rotation(int ndim, const int* ifirst, const int* ilast, const double* q, int ngq,
                              int depth, int* const* rot, int ngrot, void* stream)
{
   if (ndim != 2 && ndim != 3) return ampe_set_err(AMPE_EINVAL, "ndim must be 2 or 3");
   if (!ifirst || !ilast || !q || !rot) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (ngq < 1 || ngrot < 1) return ampe_set_err(AMPE_EINVAL, "quat_symm_rotation needs ghost width >= 1");
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV qv = view(q, b, -1, ngq);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 1, L, H);
      int rc = launch_symm_rotation(depth, a, L, H, qv, view(rot[a], b, a, ngrot), ST(stream));
      if (rc) return rc;
   }
   return AMPE_OK;
}

/* QUAT_FUNDAMENTAL (QuatFort.h:331; quatrotation.m4:93-147): in place on the box [ifirst, ilast];
 * qlo/qhi = ghost box of the array, as the reference passes it. */
int ampe_k_quat_fundamental(int ndim, const int* ifirst, const int* ilast, double* quat, const int* qlo,
                            const int* qhi, int depth, void* stream)
{
   if (ndim != 2 && ndim != 3) return ampe_set_err(AMPE_EINVAL, "ndim must be 2 or 3");
   if (!ifirst || !ilast || !quat || !qlo || !qhi) return ampe_set_err(AMPE_EINVAL, "null argument");
   int L[3] = {0, 0, 0}, H[3] = {0, 0, 0};
   for (int d = 0; d < ndim; d++) L[d] = ifirst[d], H[d] = ilast[d];
   return launch_fundamental(depth, L, H, view_lohi(quat, ndim, qlo, qhi), ST(stream));
}

/* PROJECT2D / PROJECT3D (QuatFort.h:883, 1022; 3d/quatfacops.m4:1022-1081) */
int ampe_k_project(int ndim, const int* lo, const int* hi, int depth, const double* q, const int* qlo,
                   const int* qhi, double* corr, const int* clo, const int* chi, double* err, const int* elo,
                   const int* ehi, void* stream)
{
   if (ndim != 2 && ndim != 3) return ampe_set_err(AMPE_EINVAL, "ndim must be 2 or 3");
   if (!lo || !hi || !q || !corr || !err || !qlo || !qhi || !clo || !chi || !elo || !ehi)
      return ampe_set_err(AMPE_EINVAL, "null argument");
   if (depth < 1 || depth > 4) return ampe_set_err(AMPE_EINVAL, "project: depth must be 1..4");
   int L[3] = {0, 0, 0}, H[3] = {0, 0, 0};
   for (int d = 0; d < ndim; d++) L[d] = lo[d], H[d] = hi[d];
   const int e0 = H[0] - L[0] + 1, e1 = H[1] - L[1] + 1, e2 = H[2] - L[2] + 1;
   if (e0 <= 0 || e1 <= 0 || e2 <= 0) return AMPE_OK;
   const long long total = (long long)e0 * e1 * e2;
   const CV qv = view_lohi(q, ndim, qlo, qhi);
   const DV cv = view_lohi(corr, ndim, clo, chi), ev = view_lohi(err, ndim, elo, ehi);
   const int nb = blocks_for(total);
   switch (depth) {
      case 1: project_box_kernel<1><<<nb, 256, 0, ST(stream)>>>(L[0], L[1], L[2], e0, e1, total, qv, cv, ev); break;
      case 2: project_box_kernel<2><<<nb, 256, 0, ST(stream)>>>(L[0], L[1], L[2], e0, e1, total, qv, cv, ev); break;
      case 3: project_box_kernel<3><<<nb, 256, 0, ST(stream)>>>(L[0], L[1], L[2], e0, e1, total, qv, cv, ev); break;
      default: project_box_kernel<4><<<nb, 256, 0, ST(stream)>>>(L[0], L[1], L[2], e0, e1, total, qv, cv, ev); break;
   }
   return check_last("project");
}

/* QuatModel::computeSymmetryRotations (QuatModel.cc:4978-5055) on the context's periodic level:
 * rotation index of every lower face from y->quat, kept in the context (the array
 * ampe_rhs_set_symmetry_rotations fills); the previous indices seed the search, 0 at creation. */
int ampe_rhs_compute_symmetry_rotations(ampe_rhs_ctx* c, const ampe_rhs_fields* y, void* stream)
{
   if (!c || !y) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (!c->p.symm) return ampe_set_err(AMPE_EINVAL, "context is not symmetry aware");
   if (!y->quat) return ampe_set_err(AMPE_EINVAL, "quat missing");
   // slab rank with neighbours: the caller has exchanged the ghost planes of y (ampe_halo_push / ampe_halo_wait, or
   // ampe_rhs_compute_symmetry_rotations_slab, which does both)
   const bool slab = c->have_halo && c->halo_lo.quat;
   if (c->cfg.nranks > 1 && !slab)
      return ampe_set_err(AMPE_EINVAL, "several ranks: use ampe_rhs_compute_symmetry_rotations_slab");
   int rc = ensure_table();
   if (rc) return rc;
   const Params& p = c->p;
   cudaStream_t st = (cudaStream_t)stream;
   const long long pl = c->plane;
   const int ng = c->ng, ns = c->ns;
   int* r[3];
   for (int d = 0; d < 3; d++) r[d] = c->iq[d] ? c->iq[d] + (long long)ng * pl : nullptr;
   const SymmConst sc = symm_const();
   const int nb = blocks_for(c->ncell);
   const int n2 = p.ndim == 3 ? p.n[2] : 1;
#define LAUNCH(Q, ND) \
   symm_rotation_ctx_kernel<Q, ND><<<nb, 256, 0, st>>>(p.n[0], p.n[1], n2, y->quat, r[0], r[1], r[2], sc, \
                                                       slab ? c->halo_lo.quat : nullptr, (long long)ng * pl, ng)
   if (p.ndim == 2) {
      if (p.qlen == 4) LAUNCH(4, 2);
      else if (p.qlen == 2) LAUNCH(2, 2);
      else if (p.qlen == 1) LAUNCH(1, 2);
      else return ampe_set_err(AMPE_EINVAL, "qlen must be 1, 2 or 4");
   } else {
      if (p.qlen == 4) LAUNCH(4, 3);
      else if (p.qlen == 2) LAUNCH(2, 3);
      else if (p.qlen == 1) LAUNCH(1, 3);
      else return ampe_set_err(AMPE_EINVAL, "qlen must be 1, 2 or 4");
   }
#undef LAUNCH
   rc = check_last("compute_symmetry_rotations");
   if (rc) return rc;
   if (slab) return AMPE_OK;  // the ghost planes of the indices come from the neighbours (halo.cu)
   // ghost planes along the slab axis: periodic images of the interior planes
   for (int d = 0; d < p.ndim; d++) {
      int* base = c->iq[d];
      CUDA_OKS(cudaMemcpyAsync(base, base + (long long)ns * pl, sizeof(int) * pl * ng, cudaMemcpyDeviceToDevice, st));
      CUDA_OKS(cudaMemcpyAsync(base + (long long)(ng + ns) * pl, base + (long long)ng * pl, sizeof(int) * pl * ng,
                               cudaMemcpyDeviceToDevice, st));
   }
   return AMPE_OK;
}

/* copy of the context's rotation indices (ghost 0, one array per direction) into caller-owned
 * device arrays -- what the reference exposes as the quat_symm_rotation SideData */
int ampe_rhs_get_symmetry_rotations(ampe_rhs_ctx* c, int* const* iqrot_out, void* stream)
{
   if (!c || !iqrot_out) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (!c->p.symm) return ampe_set_err(AMPE_EINVAL, "context is not symmetry aware");
   for (int d = 0; d < c->p.ndim; d++) {
      if (!iqrot_out[d]) return ampe_set_err(AMPE_EINVAL, "null output array");
      CUDA_OKS(cudaMemcpyAsync(iqrot_out[d], c->iq[d] + (long long)c->ng * c->plane, sizeof(int) * c->ncell,
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
   }
   return AMPE_OK;
}

/* QuatModel::makeQuatFundamental (QuatModel.cc:5059-5104): y->quat in place */
int ampe_quat_fundamental(ampe_rhs_ctx* c, const ampe_rhs_fields* y, void* stream)
{
   if (!c || !y) return ampe_set_err(AMPE_EINVAL, "null argument");
   const Params& p = c->p;
   if (p.qlen < 1) return AMPE_OK;
   if (!y->quat) return ampe_set_err(AMPE_EINVAL, "quat missing");
   int L[3] = {0, 0, 0}, H[3] = {p.n[0] - 1, p.n[1] - 1, (p.ndim == 3 ? p.n[2] : 1) - 1};
   return launch_fundamental(p.qlen, L, H, view_lohi(y->quat, p.ndim, L, H), (cudaStream_t)stream);
}

/* QuatIntegrator::applyProjection(time, y, corr, epsProj, err) (QuatIntegrator.cc:3911-3962): every
 * evolved component of corr is zeroed, then (qlen > 1 and the orientation is evolved) the quaternion
 * part is projected onto the unit sphere: y + corr has |q| = 1 and err loses its component along q. */
int ampe_apply_projection(ampe_rhs_ctx* c, const ampe_rhs_fields* y, const ampe_rhs_fields* corr,
                          const ampe_rhs_fields* err, void* stream)
{
   if (!c || !y || !corr || !err) return ampe_set_err(AMPE_EINVAL, "null argument");
   const Params& p = c->p;
   cudaStream_t st = (cudaStream_t)stream;
   const size_t nb = sizeof(double) * (size_t)c->ncell;
   if (p.with_phase) {
      if (!corr->phase) return ampe_set_err(AMPE_EINVAL, "corr: an evolved component is NULL");
      CUDA_OKS(cudaMemsetAsync(corr->phase, 0, nb, st));
   }
   if (p.with_conc) {
      if (!corr->conc) return ampe_set_err(AMPE_EINVAL, "corr: an evolved component is NULL");
      CUDA_OKS(cudaMemsetAsync(corr->conc, 0, nb, st));
   }
   if (p.with_T) {
      if (!corr->temperature) return ampe_set_err(AMPE_EINVAL, "corr: an evolved component is NULL");
      CUDA_OKS(cudaMemsetAsync(corr->temperature, 0, nb, st));
   }
   if (p.evolve_quat) {
      if (!y->quat || !corr->quat || !err->quat) return ampe_set_err(AMPE_EINVAL, "quat component is NULL");
      if (p.qlen > 1) {
         int L[3] = {0, 0, 0}, H[3] = {p.n[0] - 1, p.n[1] - 1, (p.ndim == 3 ? p.n[2] : 1) - 1};
         return ampe_k_project(p.ndim, L, H, p.qlen, y->quat, L, H, corr->quat, L, H, err->quat, L, H, stream);
      }
      CUDA_OKS(cudaMemsetAsync(corr->quat, 0, nb * p.qlen, st));
   }
   return AMPE_OK;
}

}  // extern "C"
