// Device-resident pieces around the RHS evaluation, behind the C ABI
// of include/ampe_b200.h:
//   * N_Vector operations CVODE needs on the solution vector (samrai/Sundials_SAMRAIVector.cc:
//     linearSum, scale, dotWith, weightedRMSNorm, maxNorm), over the evolved components;
//   * QuatModel::normalizeQuat (QuatModel.cc:4222-4262);
//   * scalar energy diagnostics (QuatModel::evaluateEnergy, quatenergy.m4) -- reduction of the
//     per-block partial sums written by energy_tile_kernel, and the PFHub-1a Cahn-Hilliard energy;
//   * a fixed-step explicit integrator (forward Euler / Heun) that keeps y on the device between
//     evaluations -- the minimal stand-in for the CVODE loop of QuatIntegrator::Advance used by
//     the trajectory parity tests.
// All bandwidth-trivial grid-stride kernels; deterministic two-stage reductions.
#include <cstring>

#include "ctx_internal.h"

using namespace ampe;

#define CUDA_OKV(call)                                                                       \
   do {                                                                                      \
      cudaError_t e_ = (call);                                                               \
      if (e_ != cudaSuccess)                                                                 \
         return ampe_set_err(AMPE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
   } while (0)

namespace {

constexpr int VT = 256;
constexpr int RED_BLOCKS = 592;  // 4 per SM on B200

int grid_for(long long n)
{
   long long b = (n + VT - 1) / VT;
   if (b > 148LL * 16) b = 148LL * 16;
   return (int)(b < 1 ? 1 : b);
}

// the evolved components of the solution vector (createSolutionvector, QuatIntegrator.cc:1623-1674)
struct Comp {
   int n = 0;
   const double* x[4];
   const double* y[4];
   double* z[4];
   long long len[4];
};

int components(const ampe_rhs_ctx* c, const ampe_rhs_fields* x, const ampe_rhs_fields* y,
               const ampe_rhs_fields* z, Comp& o)
{
   const Params& p = c->p;
   auto add = [&](bool on, double* ampe_rhs_fields::*m, long long len) -> int {
      if (!on) return 0;
      if ((x && !(x->*m)) || (y && !(y->*m)) || (z && !(z->*m)))
         return ampe_set_err(AMPE_EINVAL, "vector operation: an evolved component is NULL");
      o.x[o.n] = x ? x->*m : nullptr;
      o.y[o.n] = y ? y->*m : nullptr;
      o.z[o.n] = z ? z->*m : nullptr;
      o.len[o.n++] = len;
      return 0;
   };
   int rc = add(p.with_phase, &ampe_rhs_fields::phase, c->ncell);
   if (!rc) rc = add(p.evolve_quat, &ampe_rhs_fields::quat, c->ncell * p.qlen);
   if (!rc) rc = add(p.with_conc, &ampe_rhs_fields::conc, c->ncell);
   if (!rc) rc = add(p.with_T, &ampe_rhs_fields::temperature, c->ncell);
   return rc;
}

__global__ void linear_sum_kernel(double a, const double* x, double b, const double* y, double* z, long long n)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      z[i] = a * x[i] + b * y[i];
}
__global__ void axpy_kernel(double a, const double* x, double* z, long long n)  // z += a x
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      z[i] = z[i] + a * x[i];
}
__global__ void scale_kernel(double a, const double* x, double* z, long long n)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      z[i] = a * x[i];
}

// mode 0: sum x*y   1: sum (x*w)^2   2: max |x|
template <int MODE>
__global__ void reduce_kernel(const double* x, const double* y, long long n, double* partial)
{
   double acc = 0.0;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      if (MODE == 0) acc += x[i] * y[i];
      if (MODE == 1) {
         const double t = x[i] * y[i];
         acc += t * t;
      }
      if (MODE == 2) acc = fmax(acc, fabs(x[i]));
   }
   __shared__ double red[VT / 32];
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_down_sync(0xffffffffu, acc, o);
      acc = (MODE == 2) ? fmax(acc, t) : acc + t;
   }
   if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
   __syncthreads();
   if (threadIdx.x == 0) {
      double r = red[0];
      for (int w = 1; w < VT / 32; w++) r = (MODE == 2) ? fmax(r, red[w]) : r + red[w];
      partial[blockIdx.x] = r;
   }
}
// out[v] (+)= reduction of partial[b*nvals + v] over b, fixed order (one block, one warp per value)
template <int MODE>
__global__ void final_reduce_kernel(const double* partial, long long nblocks, int nvals, double* out, int accumulate)
{
   const int v = threadIdx.x / 32, lane = threadIdx.x % 32;
   if (v >= nvals) return;
   double acc = 0.0;
   for (long long b = lane; b < nblocks; b += 32) {
      const double t = partial[b * nvals + v];
      acc = (MODE == 2) ? fmax(acc, t) : acc + t;
   }
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_down_sync(0xffffffffu, acc, o);
      acc = (MODE == 2) ? fmax(acc, t) : acc + t;
   }
   if (lane == 0) {
      if (accumulate) acc = (MODE == 2) ? fmax(acc, out[v]) : acc + out[v];
      out[v] = acc;
   }
}

// sum w^2 x y : the inner product of CVODE's scaled Krylov solver (SPGMR with s1 = s2 = ewt)
__global__ void wdot_kernel(const double* x, const double* y, const double* w, long long n, double* partial)
{
   double acc = 0.0;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const double a = x[i] * w[i], b = y[i] * w[i];
      acc += a * b;
   }
   __shared__ double red[VT / 32];
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
   if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
   __syncthreads();
   if (threadIdx.x == 0) {
      double r = red[0];
      for (int k = 1; k < VT / 32; k++) r += red[k];
      partial[blockIdx.x] = r;
   }
}
// CVODE error weights (cvEwtSetSS): w = 1 / (rtol |y| + atol)
__global__ void ewt_kernel(const double* y, double rtol, double atol, double* w, long long n)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      w[i] = 1.0 / (rtol * fabs(y[i]) + atol);
}

// QuatModel::normalizeQuat (QuatModel.cc:4237-4262): q *= 1/sqrt(sum q^2), per cell
template <int Q>
__global__ void normalize_quat_kernel(double* q, long long ncell)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) {
      double v[Q], n2 = 0.0;
#pragma unroll
      for (int m = 0; m < Q; m++) {
         v[m] = q[i + m * ncell];
         n2 = n2 + v[m] * v[m];
      }
      const double inv = 1.0 / sqrt(n2);
#pragma unroll
      for (int m = 0; m < Q; m++) q[i + m * ncell] = v[m] * inv;
   }
}

// PFHub 1a free energy of the Cahn-Hilliard model (no evaluator in the reference; the functional
// whose variational derivative is the mu of add_cahnhilliarddoublewell_flux,
// 2d/concentrationrhs.m4:85-137): w (c-ca)^2 (cb-c)^2 + kappa/2 |grad c|^2, the gradient as the
// mean of the squared face gradients.  partial[b*6 + {0 total, 1 gradient, 4 well}]
template <int ND>
__global__ void ch_energy_kernel(const __grid_constant__ ChArgs A, double* partial)
{
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;
   const long long total = plane * ns;
   auto at = [&](int i, int j, int k) -> double {
      i = (i < 0) ? i + n0 : ((i >= n0) ? i - n0 : i);
      int sl;
      long long inplane;
      if (ND == 3) {
         j = (j < 0) ? j + n1 : ((j >= n1) ? j - n1 : j);
         sl = k;
         inplane = i + (long long)n0 * j;
      } else {
         sl = j;
         inplane = i;
      }
      if (sl < 0) return A.conc.lo[(long long)(sl + 2) * plane + inplane];
      if (sl >= ns) return A.conc.hi[(long long)(sl - ns) * plane + inplane];
      return A.conc.base[(long long)sl * plane + inplane];
   };
   double weight = p.h[0] * p.h[1];
   if (ND == 3) weight = weight * p.h[2];
   double aw = 0.0, ag = 0.0;
   for (long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x; cell < total;
        cell += (long long)gridDim.x * blockDim.x) {
      const int i = (int)(cell % n0);
      const int j = (int)((cell / n0) % n1);
      const int k = (int)(cell / ((long long)n0 * n1));
      const double c = at(i, j, k);
      const double w = p.ch_well_scale * (c - p.ch_ca) * (c - p.ch_ca) * (p.ch_cb - c) * (p.ch_cb - c);
      double g2 = 0.0;
#pragma unroll
      for (int a = 0; a < ND; a++) {
         const double gl = (c - at(i - (a == 0), j - (a == 1), k - (a == 2))) * p.dinv[a];
         const double gu = (at(i + (a == 0), j + (a == 1), k + (a == 2)) - c) * p.dinv[a];
         g2 = g2 + 0.5 * (gl * gl + gu * gu);
      }
      aw += w * weight;
      ag += (0.5 * p.ch_kappa * g2) * weight;
   }
   __shared__ double red[2][VT / 32];
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      aw += __shfl_down_sync(0xffffffffu, aw, o);
      ag += __shfl_down_sync(0xffffffffu, ag, o);
   }
   if (threadIdx.x % 32 == 0) red[0][threadIdx.x / 32] = aw, red[1][threadIdx.x / 32] = ag;
   __syncthreads();
   if (threadIdx.x == 0) {
      double w = 0.0, g = 0.0;
      for (int t = 0; t < VT / 32; t++) w += red[0][t], g += red[1][t];
      double* o = partial + (long long)blockIdx.x * 6;
      o[0] = w + g, o[1] = g, o[2] = 0.0, o[3] = 0.0, o[4] = w, o[5] = 0.0;
   }
}

// ---- scalar diagnostics (QuatModel::printScalarDiagnostics, QuatModel.cc:2543-2690) ---------------
// one pass over phi, c, T: per block sum |phi|, sum phi, sum c, sum |c phi|, sum T, max c, max T, min T
__global__ void scalar_diag_kernel(const double* phi, const double* conc, const double* T, long long n,
                                   double* partial)
{
   double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, mc = -1.0e300, mt = -1.0e300, nt = 1.0e300;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      const double p = phi ? phi[i] : 1.0, cc = conc ? conc[i] : 0.0;
      a0 += fabs(p);
      a1 += p;
      a2 += cc;
      a3 += fabs(cc * p);
      mc = fmax(mc, cc);
      if (T) {
         const double t = T[i];
         a4 += t;
         mt = fmax(mt, t);
         nt = fmin(nt, t);
      }
   }
   __shared__ double red[8][VT / 32];
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_down_sync(0xffffffffu, a0, o);
      a1 += __shfl_down_sync(0xffffffffu, a1, o);
      a2 += __shfl_down_sync(0xffffffffu, a2, o);
      a3 += __shfl_down_sync(0xffffffffu, a3, o);
      a4 += __shfl_down_sync(0xffffffffu, a4, o);
      mc = fmax(mc, __shfl_down_sync(0xffffffffu, mc, o));
      mt = fmax(mt, __shfl_down_sync(0xffffffffu, mt, o));
      nt = fmin(nt, __shfl_down_sync(0xffffffffu, nt, o));
   }
   if (threadIdx.x % 32 == 0) {
      const int w = threadIdx.x / 32;
      red[0][w] = a0, red[1][w] = a1, red[2][w] = a2, red[3][w] = a3, red[4][w] = a4;
      red[5][w] = mc, red[6][w] = mt, red[7][w] = nt;
   }
   __syncthreads();
   if (threadIdx.x < 8) {
      const int v = threadIdx.x;
      double r = red[v][0];
      for (int w = 1; w < VT / 32; w++) r = v < 5 ? r + red[v][w] : (v == 7 ? fmin(r, red[v][w]) : fmax(r, red[v][w]));
      partial[(long long)blockIdx.x * 8 + v] = r;
   }
}
// lane v < 8 folds value v over the blocks in block order (deterministic)
__global__ void scalar_diag_final_kernel(const double* partial, long long nblocks, double* out)
{
   const int v = threadIdx.x;
   if (v >= 8) return;
   double r = partial[v];
   for (long long b = 1; b < nblocks; b++) {
      const double t = partial[b * 8 + v];
      r = v < 5 ? r + t : (v == 7 ? fmin(r, t) : fmax(r, t));
   }
   out[v] = r;
}

}  // namespace

int ampe_ensure_scratch(ampe_rhs_ctx* c, long long nblocks)
{
   if (nblocks > c->partials_cap) {
      cudaFree(c->partials);
      c->partials = nullptr;
      CUDA_OKV(cudaMalloc(&c->partials, (size_t)nblocks * 8 * sizeof(double)));
      c->partials_cap = nblocks;
   }
   if (!c->red_out) CUDA_OKV(cudaMalloc(&c->red_out, 8 * sizeof(double)));
   return AMPE_OK;
}

namespace {

template <int MODE>
int reduce_components(ampe_rhs_ctx* c, const Comp& v, double* host_out, cudaStream_t st)
{
   int rc = ampe_ensure_scratch(c, RED_BLOCKS);
   if (rc) return rc;
   for (int n = 0; n < v.n; n++) {
      const int blocks = (int)((v.len[n] + VT - 1) / VT < RED_BLOCKS ? (v.len[n] + VT - 1) / VT : RED_BLOCKS);
      reduce_kernel<MODE><<<blocks, VT, 0, st>>>(v.x[n], v.y[n], v.len[n], c->partials);
      final_reduce_kernel<MODE><<<1, 32, 0, st>>>(c->partials, blocks, 1, c->red_out, n > 0);
   }
   CUDA_OKV(cudaGetLastError());
   CUDA_OKV(cudaMemcpyAsync(host_out, c->red_out, sizeof(double), cudaMemcpyDeviceToHost, st));
   CUDA_OKV(cudaStreamSynchronize(st));
   return AMPE_OK;
}

}  // namespace

// z = a x + b y
extern "C" int ampe_vec_linear_sum(ampe_rhs_ctx* c, double a, const ampe_rhs_fields* x, double b,
                                   const ampe_rhs_fields* y, const ampe_rhs_fields* z, void* stream)
{
   if (!c || !x || !y || !z) return ampe_set_err(AMPE_EINVAL, "null argument");
   Comp v;
   int rc = components(c, x, y, z, v);
   if (rc) return rc;
   for (int n = 0; n < v.n; n++)
      linear_sum_kernel<<<grid_for(v.len[n]), VT, 0, (cudaStream_t)stream>>>(a, v.x[n], b, v.y[n], v.z[n], v.len[n]);
   CUDA_OKV(cudaGetLastError());
   return AMPE_OK;
}

// z = a x
extern "C" int ampe_vec_scale(ampe_rhs_ctx* c, double a, const ampe_rhs_fields* x,
                              const ampe_rhs_fields* z, void* stream)
{
   if (!c || !x || !z) return ampe_set_err(AMPE_EINVAL, "null argument");
   Comp v;
   int rc = components(c, x, nullptr, z, v);
   if (rc) return rc;
   for (int n = 0; n < v.n; n++)
      scale_kernel<<<grid_for(v.len[n]), VT, 0, (cudaStream_t)stream>>>(a, v.x[n], v.z[n], v.len[n]);
   CUDA_OKV(cudaGetLastError());
   return AMPE_OK;
}

extern "C" int ampe_vec_dot(ampe_rhs_ctx* c, const ampe_rhs_fields* x, const ampe_rhs_fields* y,
                            double* result, void* stream)
{
   if (!c || !x || !y || !result) return ampe_set_err(AMPE_EINVAL, "null argument");
   Comp v;
   int rc = components(c, x, y, nullptr, v);
   if (rc) return rc;
   return reduce_components<0>(c, v, result, (cudaStream_t)stream);
}

// sqrt( sum (x_i w_i)^2 / N )
extern "C" int ampe_vec_wrms_norm(ampe_rhs_ctx* c, const ampe_rhs_fields* x, const ampe_rhs_fields* w,
                                  double* result, void* stream)
{
   if (!c || !x || !w || !result) return ampe_set_err(AMPE_EINVAL, "null argument");
   Comp v;
   int rc = components(c, x, w, nullptr, v);
   if (rc) return rc;
   rc = reduce_components<1>(c, v, result, (cudaStream_t)stream);
   if (rc) return rc;
   long long n = 0;
   for (int k = 0; k < v.n; k++) n += v.len[k];
   *result = sqrt(*result / (double)n);
   return AMPE_OK;
}

extern "C" int ampe_vec_max_norm(ampe_rhs_ctx* c, const ampe_rhs_fields* x, double* result, void* stream)
{
   if (!c || !x || !result) return ampe_set_err(AMPE_EINVAL, "null argument");
   Comp v;
   int rc = components(c, x, x, nullptr, v);
   if (rc) return rc;
   return reduce_components<2>(c, v, result, (cudaStream_t)stream);
}

// sum over the evolved components of (w x)(w y)
extern "C" int ampe_vec_wdot(ampe_rhs_ctx* c, const ampe_rhs_fields* x, const ampe_rhs_fields* y,
                             const ampe_rhs_fields* w, double* result, void* stream)
{
   if (!c || !x || !y || !w || !result) return ampe_set_err(AMPE_EINVAL, "null argument");
   Comp v, vw;
   int rc = components(c, x, y, nullptr, v);
   if (!rc) rc = components(c, w, nullptr, nullptr, vw);
   if (!rc) rc = ampe_ensure_scratch(c, RED_BLOCKS);
   if (rc) return rc;
   cudaStream_t st = (cudaStream_t)stream;
   for (int n = 0; n < v.n; n++) {
      const int blocks = (int)((v.len[n] + VT - 1) / VT < RED_BLOCKS ? (v.len[n] + VT - 1) / VT : RED_BLOCKS);
      wdot_kernel<<<blocks, VT, 0, st>>>(v.x[n], v.y[n], vw.x[n], v.len[n], c->partials);
      final_reduce_kernel<0><<<1, 32, 0, st>>>(c->partials, blocks, 1, c->red_out, n > 0);
   }
   CUDA_OKV(cudaGetLastError());
   CUDA_OKV(cudaMemcpyAsync(result, c->red_out, sizeof(double), cudaMemcpyDeviceToHost, st));
   CUDA_OKV(cudaStreamSynchronize(st));
   return AMPE_OK;
}

// w = 1 / (rtol |y| + atol) on the evolved components (CVODE cvEwtSetSS)
extern "C" int ampe_vec_error_weights(ampe_rhs_ctx* c, const ampe_rhs_fields* y, double rtol, double atol,
                                      const ampe_rhs_fields* w, void* stream)
{
   if (!c || !y || !w) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (!(rtol >= 0.0) || !(atol >= 0.0) || (rtol == 0.0 && atol == 0.0))
      return ampe_set_err(AMPE_EINVAL, "error weights: rtol, atol >= 0 and not both 0 (CVODESolver.cc:147-151)");
   Comp v;
   int rc = components(c, y, nullptr, w, v);
   if (rc) return rc;
   for (int n = 0; n < v.n; n++)
      ewt_kernel<<<grid_for(v.len[n]), VT, 0, (cudaStream_t)stream>>>(v.x[n], rtol, atol, v.z[n], v.len[n]);
   CUDA_OKV(cudaGetLastError());
   return AMPE_OK;
}

extern "C" int ampe_normalize_quat(ampe_rhs_ctx* c, const ampe_rhs_fields* y, void* stream)
{
   if (!c || !y) return ampe_set_err(AMPE_EINVAL, "null argument");
   const Params& p = c->p;
   if (p.qlen <= 1) return AMPE_OK;
   if (!y->quat) return ampe_set_err(AMPE_EINVAL, "quat missing");
   cudaStream_t st = (cudaStream_t)stream;
   if (p.qlen == 2)
      normalize_quat_kernel<2><<<grid_for(c->ncell), VT, 0, st>>>(y->quat, c->ncell);
   else
      normalize_quat_kernel<4><<<grid_for(c->ncell), VT, 0, st>>>(y->quat, c->ncell);
   CUDA_OKV(cudaGetLastError());
   return AMPE_OK;
}

// out[8]: total, phase interface, orientational, q interface, double well, bulk free energy, 0, 0
// (this rank's cells only: a multi-rank caller sums the ranks)
extern "C" int ampe_energy_eval(ampe_rhs_ctx* c, const ampe_rhs_fields* y, double* out, void* stream)
{
   if (!c || !y || !out) return ampe_set_err(AMPE_EINVAL, "null argument");
   const Params& p = c->p;
   cudaStream_t st = (cudaStream_t)stream;
   for (int n = 0; n < 8; n++) out[n] = 0.0;
   long long nblocks = 0;
   if (p.conc_form == AMPE_CONC_CAHN_HILLIARD) {
      if (!y->conc) return ampe_set_err(AMPE_EINVAL, "conc missing");
      nblocks = RED_BLOCKS;
      int rc = ampe_ensure_scratch(c, nblocks);
      if (rc) return rc;
      ChArgs A;
      A.p = p;
      A.conc.base = y->conc;
      A.conc.comp = c->ncell;
      if (c->have_halo && c->halo_lo.conc && c->halo_hi.conc) {
         A.conc.lo = c->halo_lo.conc;
         A.conc.hi = c->halo_hi.conc;
      } else {
         A.conc.lo = y->conc + (long long)(c->ns - c->ng) * c->plane;
         A.conc.hi = y->conc;
      }
      A.conc.hcomp = 0;
      A.out_c = nullptr;
      A.s_begin = 0;
      A.s_end = c->ns;
      if (p.ndim == 2)
         ch_energy_kernel<2><<<(int)nblocks, VT, 0, st>>>(A, c->partials);
      else
         ch_energy_kernel<3><<<(int)nblocks, VT, 0, st>>>(A, c->partials);
   } else {
      if (!p.with_phase) return AMPE_OK;  // QuatModel.cc:4953: no evaluator without a phase field
      if (p.ndim == 3 && p.nu > 0.0)
         return ampe_set_err(AMPE_EINVAL, "3D anisotropic interface energy is not on the path");
      int rc = ampe_ensure_scratch(c, 1);
      if (rc) return rc;
      rc = ampe_launch_energy(c, y, st, &nblocks);
      if (rc) return rc;
   }
   final_reduce_kernel<0><<<1, 32 * 6, 0, st>>>(c->partials, nblocks, 6, c->red_out, 0);
   CUDA_OKV(cudaGetLastError());
   CUDA_OKV(cudaMemcpyAsync(out, c->red_out, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
   CUDA_OKV(cudaStreamSynchronize(st));
   return AMPE_OK;
}

// Fixed-step explicit integration of nsteps steps, y updated in place on the device.
//   scheme 0: forward Euler   y += dt f(y)
//   scheme 1: Heun (RK2)      y* = y + dt f(y);  y += dt/2 (f(y) + f(y*))
// After every step the quaternions are renormalised (QuatModel::normalizeQuat, as
// QuatModel::Advance does after the integrator returns) and, for the CALPHAD models, the Newton
// initial guess is refreshed from the converged c_l, c_a (resetRefPhaseConcentrations,
// QuatModel.cc:5218-5231).  work1/work2: caller-owned vectors shaped like y (work2 only for Heun).
// h != NULL: slab rank with neighbours -- every evaluation exchanges the ghost planes (ampe_rhs_eval_slab); the
// updates of y, the normalisation and the Newton reference reset are local to the rank
static int integrate_fixed_impl(ampe_rhs_ctx* c, ampe_halo* h, const ampe_rhs_fields* y, const ampe_rhs_fields* work1,
                                const ampe_rhs_fields* work2, double t0, double dt, int nsteps, int scheme, void* stream)
{
   if (!c || !y || !work1) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (scheme != 0 && scheme != 1) return ampe_set_err(AMPE_EINVAL, "scheme: 0 (Euler) or 1 (Heun)");
   if (scheme == 1 && !work2) return ampe_set_err(AMPE_EINVAL, "Heun needs work2");
   if (c->have_halo && !h)
      return ampe_set_err(AMPE_EINVAL, "several ranks: use ampe_integrate_fixed_slab (or drive the exchange from the caller)");
   auto rhs_eval = [&](double tt, const ampe_rhs_fields* yy, const ampe_rhs_fields* yd) {
      return h ? ampe_rhs_eval_slab(c, h, tt, yy, yd, 0, stream) : ampe_rhs_eval(c, tt, yy, yd, 0, stream);
   };
   const Params& p = c->p;
   cudaStream_t st = (cudaStream_t)stream;
   Comp vy;
   int rc = components(c, work1, nullptr, const_cast<ampe_rhs_fields*>(y), vy);
   if (rc) return rc;
   const bool kks = p.conc_form == AMPE_CONC_KKS || p.conc_form == AMPE_CONC_EBS;
   double t = t0;
   for (int s = 0; s < nsteps; s++) {
      rc = rhs_eval(t, y, work1);
      if (rc) return rc;
      if (scheme == 0) {
         for (int n = 0; n < vy.n; n++)
            axpy_kernel<<<grid_for(vy.len[n]), VT, 0, st>>>(dt, vy.x[n], vy.z[n], vy.len[n]);
      } else {
         // work2 <- y + dt k1 (all components, the non-evolved ones copied)
         ampe_rhs_fields ystar = *work2;
         Comp vs;
         rc = components(c, y, work1, &ystar, vs);
         if (rc) return rc;
         for (int n = 0; n < vs.n; n++)
            linear_sum_kernel<<<grid_for(vs.len[n]), VT, 0, st>>>(1.0, vs.x[n], dt, vs.y[n], vs.z[n], vs.len[n]);
         if (p.qlen > 0 && !p.evolve_quat)
            CUDA_OKV(cudaMemcpyAsync(ystar.quat, y->quat, sizeof(double) * c->ncell * p.qlen,
                                     cudaMemcpyDeviceToDevice, st));
         // y += dt/2 k1 now; k2 overwrites work1 afterwards
         for (int n = 0; n < vy.n; n++)
            axpy_kernel<<<grid_for(vy.len[n]), VT, 0, st>>>(0.5 * dt, vy.x[n], vy.z[n], vy.len[n]);
         rc = rhs_eval(t + dt, &ystar, work1);
         if (rc) return rc;
         for (int n = 0; n < vy.n; n++)
            axpy_kernel<<<grid_for(vy.len[n]), VT, 0, st>>>(0.5 * dt, vy.x[n], vy.z[n], vy.len[n]);
      }
      if (p.evolve_quat) {
         rc = ampe_normalize_quat(c, y, stream);
         if (rc) return rc;
      }
      if (kks && p.free_energy == AMPE_FE_CALPHAD) {
         rc = ampe_rhs_set_ref_concentrations(c, nullptr, nullptr, stream);
         if (rc) return rc;
      }
      t += dt;
   }
   CUDA_OKV(cudaGetLastError());
   return AMPE_OK;
}

extern "C" int ampe_integrate_fixed(ampe_rhs_ctx* c, const ampe_rhs_fields* y, const ampe_rhs_fields* work1,
                                    const ampe_rhs_fields* work2, double t0, double dt, int nsteps,
                                    int scheme, void* stream)
{
   return integrate_fixed_impl(c, nullptr, y, work1, work2, t0, dt, nsteps, scheme, stream);
}
extern "C" int ampe_integrate_fixed_slab(ampe_rhs_ctx* c, ampe_halo* h, const ampe_rhs_fields* y,
                                         const ampe_rhs_fields* work1, const ampe_rhs_fields* work2, double t0,
                                         double dt, int nsteps, int scheme, void* stream)
{
   if (!h) return ampe_set_err(AMPE_EINVAL, "ampe_integrate_fixed_slab: null halo");
   return integrate_fixed_impl(c, h, y, work1, work2, t0, dt, nsteps, scheme, stream);
}

// QuatModel::printScalarDiagnostics (QuatModel.cc:2543-2690) on this rank's cells.  out[12]: domain volume,
// volume of solid (evaluateVolumeSolid: L1 norm of phi), its fraction, integral concentration
// (evaluateIntegralConcentration), max concentration, integral phase concentration
// (evaluateIntegralPhaseConcentration: L1 norm of c phi), Cex = (cphi - c0 vphi) / c0V0, min / max / average
// temperature, thermal energy (computeThermalEnergy: -L int phi + int cp T), 0
extern "C" int ampe_scalar_diagnostics(ampe_rhs_ctx* c, const ampe_rhs_fields* y, double* out, void* stream)
{
   if (!c || !y || !out) return ampe_set_err(AMPE_EINVAL, "null argument");
   const Params& p = c->p;
   const ampe_rhs_config& cfg = c->cfg;
   if ((p.with_phase && !y->phase) || (p.with_conc && !y->conc) || (p.with_T && !y->temperature))
      return ampe_set_err(AMPE_EINVAL, "scalar diagnostics: a component of y is NULL");
   cudaStream_t st = (cudaStream_t)stream;
   const int blocks = (int)((c->ncell + VT - 1) / VT < RED_BLOCKS ? (c->ncell + VT - 1) / VT : RED_BLOCKS);
   int rc = ampe_ensure_scratch(c, RED_BLOCKS);
   if (rc) return rc;
   scalar_diag_kernel<<<blocks, VT, 0, st>>>(p.with_phase ? y->phase : nullptr, p.with_conc ? y->conc : nullptr,
                                            p.with_T ? y->temperature : nullptr, c->ncell, c->partials);
   scalar_diag_final_kernel<<<1, 32, 0, st>>>(c->partials, blocks, c->red_out);
   CUDA_OKV(cudaGetLastError());
   double r[8];
   CUDA_OKV(cudaMemcpyAsync(r, c->red_out, sizeof(r), cudaMemcpyDeviceToHost, st));
   CUDA_OKV(cudaStreamSynchronize(st));
   double dv = 1.0;
   for (int d = 0; d < cfg.ndim; d++) dv *= cfg.dx[d];
   const double vol = dv * (double)c->ncell;
   for (int n = 0; n < 12; n++) out[n] = 0.0;
   out[0] = vol;
   const double vphi = p.with_phase ? r[0] * dv : vol;  // QuatModel.cc:2606
   out[1] = vphi;
   out[2] = vphi / vol;
   if (p.with_conc) {
      const double c0V0 = r[2] * dv, cphi = r[3] * dv, c0 = c0V0 / vol;
      out[3] = c0V0;
      out[4] = r[5];
      out[5] = cphi;
      out[6] = (cphi - c0 * vphi) / c0V0;
   }
   if (p.with_T) {
      out[7] = r[7], out[8] = r[6], out[9] = r[4] * dv / vol;
      out[10] = (p.with_phase ? -1. * cfg.latent_heat * (r[1] * dv) : 0.0) + cfg.cp * (r[4] * dv);
   } else {
      out[7] = out[8] = out[9] = cfg.T_uniform;
   }
   return AMPE_OK;
}
