// Block preconditioners of the CVODE Newton-Krylov loop on the device, behind
// the C ABI of include/ampe_b200_precond.h:
//   ampe_mg_set_elliptic   EllipticFACOps::setM / setC / setD* + PoissonSpecifications
//                          (PhaseFACOps.cc:33-51, ConcFACOps.cc:19-50, QuatIntegrator.cc:3340-3346)
//   ampe_mg_set_quat       QuatFACOps::setOperatorCoefficients (QuatFACOps.cc:735-818) ->
//                          QuatLevelSolver::setMatrixCoefficients (set_j_ij / set_stencil,
//                          2d/quatlevelsolver.m4:9-118)
//   ampe_mg_solve          EllipticFACSolver::solveSystem / QuatSysSolver::solveSystem
//                          (QuatSysSolver.cc:296-347, including divide/multiplyMobilitySqrt)
//   ampe_mg_apply          the operator itself (efo_compfluxvardc + efo_compresvarsca,
//                          2d/ellipticfacops.m4:16-56, 346-393)
//   ampe_k_phasefacops_setc  PhaseFACOps::setCOnPatchPrivate (PhaseFACOps.cc:100-186)
// The reference solves its single level with hypre PFMG (third party, not in its tree).  Here:
// geometric multigrid V-cycles, every level resident in HBM, one thread per cell; the smoother is
// the reference's red-black Gauss-Seidel update.  Per-cell arithmetic: mg_cell.h.  A FIXED number
// of cycles from a zero initial guess makes the solve a fixed linear operator, which is what right-
// preconditioned GMRES needs, and needs no host synchronisation.
// All kernels are HBM-bound sweeps (a half sweep reads u, f, c, m, the ND face arrays and writes u:
// (5 + ND) * 8 bytes per updated cell before cache reuse of the neighbours, less where a coefficient is a
// constant of the block: constants are not stored -- phase 4, temperature 3, composition 3 + ND arrays).
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ampe_b200_precond.h"
#include "ctx_internal.h"
#include "mg_cell.h"

using ampe_mg_cell::Level;

#define CUDA_OKM(call)                                                                       \
   do {                                                                                      \
      cudaError_t e_ = (call);                                                               \
      if (e_ != cudaSuccess)                                                                 \
         return ampe_set_err(AMPE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
   } while (0)

struct ampe_mg {
   int ndim = 2;
   int n[3] = {1, 1, 1};
   double inv_h2[3] = {0, 0, 0};
   bool with_s = false;          // quaternion block: column multiplier present
   int ncomp = 1;                // components solved together with the one matrix (quaternion block: qlen)
   bool coefficients_set = false;
   std::vector<Level> levels;
   std::vector<double*> blocks;  // work arrays u, f, r (and s): one allocation per level
   // coefficient arrays, allocated the first time a set_* call needs them as ARRAYS; a coefficient that is
   // a constant of the block stays unallocated (null pointer in the Level, value in *_const)
   std::vector<double*> own_c, own_m;
   std::vector<double*> own_d[3];
   std::vector<char> two_colour; // level can be red-black coloured (all extents even)
   int pre = 1, post = 1, coarse = 8;
   int zero_slope[3] = {0, 0, 0};  // homogeneous Neumann boundary per direction (ampe_mg_set_zero_slope)
   int launches = 0;
   int tail_level = -1;  // first level handled by the one-block tail kernel (-1: none)
   // fused red-black sweep (mg_rb_tile_pass): per level the second u array of the ping-pong and the tile
   // shape; t[0] == 0: the level uses the two colour half-sweeps
   // alt_u[l] is the partner of Level.u and is SWAPPED with it after every sweep, so after an odd number of
   // sweeps it points into blocks[l]; alt_owned[l] keeps the allocation made for it and is what destroy frees
   std::vector<double*> alt_u, alt_owned;
   std::vector<ampe_mg_cell::TileShape> tile;
   // AMPE_B200_MG_GRAPH=1 (opt-in): the launches of one solve captured once per (rhs, soln, cycles, form)
   bool use_graph = false;
   cudaGraphExec_t graph_exec = nullptr;
   const double* graph_rhs = nullptr;
   double* graph_soln = nullptr;
   int graph_cycles = 0, graph_symm = 0, graph_launches = 0;
};

namespace {
using namespace ampe_mg_cell;

constexpr int MT = 256;
constexpr double JACOBI_OMEGA = 0.8;

struct P3 {
   const double* a[3];
};

long long cells(const Level& L) { return (long long)L.n[0] * L.n[1] * L.n[2]; }
int grid_for(long long n)
{
   long long b = (n + MT - 1) / MT;
   const long long cap = 148LL * 32;  // grid-stride beyond 32 blocks per SM
   if (b > cap) b = cap;
   return (int)(b < 1 ? 1 : b);
}

__device__ __forceinline__ void decode(const Level& L, long long idx, int& i, int& j, int& k)
{
   i = (int)(idx % L.n[0]);
   const long long t = idx / L.n[0];
   j = (int)(t % L.n[1]);
   k = (int)(t / L.n[1]);
}

__global__ void mg_smooth_rb_kernel(Level L, int colour)
{
   // one thread per cell of the colour: (i + j + k) & 1 == colour; n[0] is even on these levels
   const int h0 = L.n[0] >> 1;
   const long long total = (long long)h0 * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      const int ii = (int)(idx % h0);
      const long long t = idx / h0;
      const int j = (int)(t % L.n[1]);
      const int k = (int)(t / L.n[1]);
      const int i = 2 * ii + ((j + k + colour) & 1);
      mg_smooth_cell(L, i, j, k);
   }
}
__global__ void mg_residual_kernel(Level L)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      mg_residual_cell(L, i, j, k);
   }
}
__global__ void mg_jacobi_kernel(Level L, double omega)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      mg_jacobi_cell(L, omega, i, j, k);
   }
}
__global__ void mg_restrict_kernel(Level F, Level C)
{
   const long long total = (long long)C.n[0] * C.n[1] * C.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(C, idx, i, j, k);
      mg_restrict_cell(F, C, i, j, k);
   }
}
__global__ void mg_restrict_residual_kernel(Level F, Level C)
{
   const long long total = (long long)C.n[0] * C.n[1] * C.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(C, idx, i, j, k);
      mg_restrict_residual_cell(F, C, i, j, k);
   }
}
__global__ void mg_coarsen_kernel(Level F, Level C)
{
   const long long total = (long long)C.n[0] * C.n[1] * C.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(C, idx, i, j, k);
      mg_coarsen_cell(F, C, i, j, k);
   }
}
__global__ void mg_prolong_kernel(Level C, Level F)
{
   const long long total = (long long)F.n[0] * F.n[1] * F.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(F, idx, i, j, k);
      mg_prolong_cell(C, F, i, j, k);
   }
}
__global__ void mg_apply_kernel(Level L, const double* u, double* out)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      // every component of a multi-component solver (quaternion block): one matrix, component stride L.cs
      for (int m = 0; m < L.nc; m++) out[idx + m * L.cs] = mg_apply_cell(L, u + m * L.cs, i, j, k);
   }
}
// f = rhs (divided by s: QuatFACOps::divideMobilitySqrt), u = 0
__global__ void mg_load_kernel(Level L, const double* rhs, int symmetrized)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      for (int m = 0; m < L.nc; m++) {
         const long long q = idx + m * L.cs;
         L.f[q] = (symmetrized && L.s) ? rhs[q] / L.s[idx] : rhs[q];
         L.u[q] = 0.0;
      }
   }
}
// soln = u (times s: QuatFACOps::multiplyMobilitySqrt)
__global__ void mg_store_kernel(Level L, double* soln, int symmetrized)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x)
      for (int m = 0; m < L.nc; m++) {
         const long long q = idx + m * L.cs;
         soln[q] = (symmetrized && L.s) ? L.u[q] * L.s[idx] : L.u[q];
      }
}
__global__ void mg_set_elliptic_kernel(Level L, const double* m, int ngm, const double* c, int ngc, P3 d, P3 d2,
                                       int have_d2, int ngd, double d_scale, double ih0, double ih1, double ih2)
{
   const double inv_h2[3] = {ih0, ih1, ih2};
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      // (no diffusion arrays given: a constant D, possibly stored as an array for a zero-slope boundary)
      mg_set_elliptic_cell(L, m, ngm, c, ngc, d.a[0] ? d.a : nullptr, have_d2 ? d2.a : nullptr, ngd, d_scale, inv_h2, i, j,
                           k);
   }
}
__global__ void mg_boundary_faces_kernel(Level L, int fill)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      mg_boundary_faces_cell(L, fill, i, j, k);
   }
}
__global__ void mg_set_quat_kernel(Level L, double gamma, const double* mobility, int ngm, P3 fc, int ngfc,
                                   double ih0, double ih1, double ih2)
{
   const double inv_h2[3] = {ih0, ih1, ih2};
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      mg_set_quat_cell(L, gamma, mobility, ngm, fc.a, ngfc, inv_h2, i, j, k);
   }
}
// PhaseFACOps::setCOnPatchPrivate: C = 1 + gamma m w g''(phi)
__global__ void phasefacops_setc_kernel(Level L, const double* phi, int ngphi, const double* m, int ngm,
                                        double gamma, double well_scale, char well_type, double* c, int ngc)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      int i, j, k;
      decode(L, idx, i, j, k);
      const double p = phi[mg_samrai_index(L, -1, ngphi, i, j, k)];
      const double mm = m[mg_samrai_index(L, -1, ngm, i, j, k)];
      // second_deriv_well_func (functions.f): 'd' 32 (1 + 6 phi (phi - 1)), 's' 2
      const double g2 = (well_type == 's') ? 2.0 : 32.0 * (1.0 + 6.0 * p * (p - 1.0));
      const double gamma_m = gamma * mm;
      c[mg_samrai_index(L, -1, ngc, i, j, k)] = 1.0 + gamma_m * well_scale * g2;
   }
}

// ---- the coarse tail of a V-cycle in ONE block ------------------------------------------------
// Below TAIL_CELLS cells a level is pure launch latency (7 launches of a few microseconds per level,
// 5-6 such levels below 64^2 / 16^3).  One block of 1024 threads walks all of them -- descent,
// coarsest-level sweeps, ascent -- on the same global arrays (L2 resident) with __syncthreads()
// between the phases; same per-cell functions, same order of phases, so the result is the one of
// the per-level launches bit for bit.
constexpr int TAIL_MAX_LEVELS = 8;
constexpr long long TAIL_CELLS = 4096;
constexpr int TAIL_THREADS = 1024;
struct TailLevels {
   Level L[TAIL_MAX_LEVELS];
   int two_colour[TAIL_MAX_LEVELS];
   int n;
};

__device__ void tail_smooth(const Level& L, int two_colour, int sweeps)
{
   const long long total = (long long)L.n[0] * L.n[1] * L.n[2];
   for (int s = 0; s < sweeps; s++) {
      if (two_colour) {
         const int h0 = L.n[0] >> 1;
         const long long half = (long long)h0 * L.n[1] * L.n[2];
         for (int colour = 0; colour < 2; colour++) {
            for (long long idx = threadIdx.x; idx < half; idx += blockDim.x) {
               const int ii = (int)(idx % h0);
               const long long t = idx / h0;
               const int j = (int)(t % L.n[1]);
               const int k = (int)(t / L.n[1]);
               mg_smooth_cell(L, 2 * ii + ((j + k + colour) & 1), j, k);
            }
            __syncthreads();
         }
      } else {
         for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
            int i, j, k;
            decode(L, idx, i, j, k);
            mg_residual_cell(L, i, j, k);
         }
         __syncthreads();
         for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
            int i, j, k;
            decode(L, idx, i, j, k);
            mg_jacobi_cell(L, JACOBI_OMEGA, i, j, k);
         }
         __syncthreads();
      }
   }
}

__global__ void __launch_bounds__(TAIL_THREADS) mg_tail_kernel(TailLevels T, int pre, int post, int coarse)
{
   for (int l = 0; l + 1 < T.n; l++) {
      const Level& F = T.L[l];
      const Level& C = T.L[l + 1];
      tail_smooth(F, T.two_colour[l], pre);
      const long long nf = (long long)F.n[0] * F.n[1] * F.n[2], nc = (long long)C.n[0] * C.n[1] * C.n[2];
      for (long long idx = threadIdx.x; idx < nf; idx += blockDim.x) {
         int i, j, k;
         decode(F, idx, i, j, k);
         mg_residual_cell(F, i, j, k);
      }
      __syncthreads();
      for (long long idx = threadIdx.x; idx < nc; idx += blockDim.x) {
         int i, j, k;
         decode(C, idx, i, j, k);
         mg_restrict_cell(F, C, i, j, k);
      }
      __syncthreads();
   }
   tail_smooth(T.L[T.n - 1], T.two_colour[T.n - 1], coarse);
   for (int l = T.n - 2; l >= 0; l--) {
      const Level& F = T.L[l];
      const Level& C = T.L[l + 1];
      const long long nf = (long long)F.n[0] * F.n[1] * F.n[2];
      for (long long idx = threadIdx.x; idx < nf; idx += blockDim.x) {
         int i, j, k;
         decode(F, idx, i, j, k);
         mg_prolong_cell(C, F, i, j, k);
      }
      __syncthreads();
      tail_smooth(F, T.two_colour[l], post);
   }
}

// one red-black sweep in one pass: a block per tile, the tile of u (halo two) in shared memory
__global__ void __launch_bounds__(MT) mg_rb_fused_kernel(Level L, const double* u_in, double* u_out, TileShape T, int nt0,
                                                         int nt1)
{
   extern __shared__ double mg_tile[];
   const int b = blockIdx.x;
   const int t0 = b % nt0, t1 = (b / nt0) % nt1, t2 = b / (nt0 * nt1);
   // the components share the matrix: the coefficient reads of the second and later components hit L1 / L2
   for (int m = 0; m < L.nc; m++)
      mg_rb_tile_pass(L, L.f + m * L.cs, u_in + m * L.cs, u_out + m * L.cs, mg_tile, T, t0 * T.t[0], t1 * T.t[1],
                      t2 * T.t[2], (int)threadIdx.x, (int)blockDim.x);
}

void smooth(ampe_mg* g, int l, int sweeps, cudaStream_t st)
{
   const Level L = g->levels[l];
   const long long nc = cells(L);
   for (int s = 0; s < sweeps; s++) {
      if (g->tile[l].t[0] > 0) {
         const TileShape T = g->tile[l];
         Level& Lm = g->levels[l];
         const int nt0 = Lm.n[0] / T.t[0], nt1 = Lm.n[1] / T.t[1], nt2 = Lm.n[2] / T.t[2];
         const size_t smem = sizeof(double) * (T.t[0] + 4) * (T.t[1] + 4) * (Lm.ndim == 3 ? T.t[2] + 4 : 1);
         mg_rb_fused_kernel<<<nt0 * nt1 * nt2, MT, smem, st>>>(Lm, Lm.u, g->alt_u[l], T, nt0, nt1);
         std::swap(Lm.u, g->alt_u[l]);  // the sweep wrote the other array of the ping-pong
         g->launches += 1;
         continue;
      }
      if (g->two_colour[l]) {
         mg_smooth_rb_kernel<<<grid_for(nc / 2), MT, 0, st>>>(L, 0);
         mg_smooth_rb_kernel<<<grid_for(nc / 2), MT, 0, st>>>(L, 1);
      } else {
         mg_residual_kernel<<<grid_for(nc), MT, 0, st>>>(L);
         mg_jacobi_kernel<<<grid_for(nc), MT, 0, st>>>(L, JACOBI_OMEGA);
      }
      g->launches += 2;
   }
}

void vcycle(ampe_mg* g, cudaStream_t st)
{
   const int nl = (int)g->levels.size();
   // levels [0, top) one launch per phase; levels [top, nl) inside the one-block tail kernel
   const int top = (g->tail_level >= 0) ? g->tail_level : nl - 1;
   for (int l = 0; l < top; l++) {
      smooth(g, l, g->pre, st);
      // residual and restriction in one pass (r is neither written nor re-read)
      mg_restrict_residual_kernel<<<grid_for(cells(g->levels[l + 1])), MT, 0, st>>>(g->levels[l], g->levels[l + 1]);
      g->launches += 1;
   }
   if (g->tail_level >= 0) {
      TailLevels T;
      T.n = nl - top;
      for (int l = top; l < nl; l++) {
         T.L[l - top] = g->levels[l];
         T.two_colour[l - top] = g->two_colour[l];
      }
      mg_tail_kernel<<<1, TAIL_THREADS, 0, st>>>(T, g->pre, g->post, g->coarse);
      g->launches += 1;
   } else {
      smooth(g, nl - 1, g->coarse, st);
   }
   for (int l = top - 1; l >= 0; l--) {
      mg_prolong_kernel<<<grid_for(cells(g->levels[l])), MT, 0, st>>>(g->levels[l + 1], g->levels[l]);
      g->launches += 1;
      smooth(g, l, g->post, st);
   }
}

// which coefficients of the block are arrays, which are constants: pointers and constants of every level
// (a constant C or M is the same on every level; a constant D / h^2 is divided by 4 per level, exactly
// what averaging an array of identical values gives)
int configure(ampe_mg* g, bool var_c, double c_const, bool var_m, double m_const, bool var_d, double d_const)
{
   // a captured solve carries the level descriptors (pointers AND constants) as kernel arguments
   if (g->graph_exec) cudaGraphExecDestroy(g->graph_exec), g->graph_exec = nullptr;
   const size_t nl = g->levels.size();
   auto ensure = [&](std::vector<double*>& own, size_t l) -> cudaError_t {
      if (own.size() < nl) own.resize(nl, nullptr);
      if (own[l]) return cudaSuccess;
      return cudaMalloc(&own[l], sizeof(double) * cells(g->levels[l]));
   };
   double scale = 1.0;
   for (size_t l = 0; l < nl; l++) {
      Level& L = g->levels[l];
      L.c = nullptr, L.m = nullptr;
      L.c_const = c_const, L.m_const = m_const;
      if (var_c) {
         CUDA_OKM(ensure(g->own_c, l));
         L.c = g->own_c[l];
      }
      if (var_m) {
         CUDA_OKM(ensure(g->own_m, l));
         L.m = g->own_m[l];
      }
      for (int a = 0; a < 3; a++) {
         L.d[a] = nullptr;
         L.d_const[a] = a < g->ndim ? d_const * g->inv_h2[a] * scale : 0.0;
         if (var_d && a < g->ndim) {
            CUDA_OKM(ensure(g->own_d[a], l));
            L.d[a] = g->own_d[a][l];
         }
      }
      scale *= 0.25;
   }
   return AMPE_OK;
}

int build_coarse(ampe_mg* g, cudaStream_t st)
{
   for (size_t l = 0; l + 1 < g->levels.size(); l++) {
      mg_coarsen_kernel<<<grid_for(cells(g->levels[l + 1])), MT, 0, st>>>(g->levels[l], g->levels[l + 1]);
      g->launches += 1;
   }
   CUDA_OKM(cudaGetLastError());
   g->coefficients_set = true;
   return AMPE_OK;
}

}  // namespace

extern "C" {

int ampe_mg_create_multi(int ndim, const int* n, const double* dx, int with_column_scale, int ncomp, ampe_mg** out)
{
   if (!out || !n || !dx || (ndim != 2 && ndim != 3) || ncomp < 1 || ncomp > 8)
      return ampe_set_err(AMPE_EINVAL, "ampe_mg_create: bad argument");
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
      return ampe_set_err(AMPE_ENOGPU, "ampe_mg_create: no CUDA device (there is no CPU fallback)");
   for (int d = 0; d < ndim; d++)
      if (n[d] < 1 || !(dx[d] > 0.0)) return ampe_set_err(AMPE_EINVAL, "ampe_mg_create: bad extent or spacing");
   ampe_mg* g = new ampe_mg;
   g->ndim = ndim;
   g->with_s = with_column_scale != 0;
   g->ncomp = ncomp;
   for (int d = 0; d < 3; d++) {
      g->n[d] = d < ndim ? n[d] : 1;
      g->inv_h2[d] = d < ndim ? 1.0 / (dx[d] * dx[d]) : 0.0;
   }
   // levels: halve every direction while all extents are even and the coarse extents stay >= 2
   int cur[3] = {g->n[0], g->n[1], g->n[2]};
   for (int l = 0; l < 16; l++) {
      Level L;
      L.ndim = ndim;
      for (int d = 0; d < 3; d++) L.n[d] = cur[d];
      const long long nc = (long long)cur[0] * cur[1] * cur[2];
      const int narr = (g->with_s ? 1 : 0) + 3 * ncomp;
      double* blk = nullptr;
      cudaError_t e = cudaMalloc(&blk, sizeof(double) * nc * narr);
      if (e != cudaSuccess) {
         for (double* b : g->blocks) cudaFree(b);
         delete g;
         return ampe_set_err(AMPE_ECUDA, std::string("ampe_mg_create: cudaMalloc: ") + cudaGetErrorString(e));
      }
      cudaMemset(blk, 0, sizeof(double) * nc * narr);
      g->blocks.push_back(blk);
      double* p = blk;
      L.c = L.m = L.s = nullptr;
      L.c_const = 1.0, L.m_const = 1.0;
      if (g->with_s) L.s = p, p += nc;
      for (int d = 0; d < 3; d++) L.d[d] = nullptr, L.d_const[d] = 0.0;
      L.nc = ncomp, L.cs = nc;
      L.clamp[0] = L.clamp[1] = L.clamp[2] = 0;
      L.u = p, p += nc * ncomp;
      L.f = p, p += nc * ncomp;
      L.r = p, p += nc * ncomp;
      g->levels.push_back(L);
      bool even = true;
      for (int d = 0; d < ndim; d++) even = even && (cur[d] % 2 == 0);
      g->two_colour.push_back(even ? 1 : 0);
      bool can = even;
      for (int d = 0; d < ndim; d++) can = can && (cur[d] / 2 >= 2);
      if (!can) break;
      for (int d = 0; d < ndim; d++) cur[d] /= 2;
   }
   // the coarse tail: every level from the first one with <= TAIL_CELLS cells on (AMPE_B200_MG_TAIL=0: off)
   const char* tail_env = getenv("AMPE_B200_MG_TAIL");
   if (!(tail_env && tail_env[0] == '0')) {
      const int nl = (int)g->levels.size();
      for (int l = 0; l < nl; l++)
         if (cells(g->levels[l]) <= TAIL_CELLS && nl - l <= TAIL_MAX_LEVELS) {
            g->tail_level = l;
            break;
         }
   }
   const char* graph_env = getenv("AMPE_B200_MG_GRAPH");
   g->use_graph = graph_env && graph_env[0] == '1';
   // fused red-black sweeps on the levels above the tail whose extents the tile divides (AMPE_B200_MG_FUSED=0:
   // off; off under graph capture, whose kernel arguments would freeze the ping-pong pointers)
   const char* fused_env = getenv("AMPE_B200_MG_FUSED");
   const bool fused = !(fused_env && fused_env[0] == '0') && !g->use_graph;
   g->alt_u.assign(g->levels.size(), nullptr);
   g->alt_owned.assign(g->levels.size(), nullptr);
   g->tile.assign(g->levels.size(), TileShape{{0, 0, 0}});
   for (size_t l = 0; fused && l < g->levels.size(); l++) {
      const Level& L = g->levels[l];
      if (!g->two_colour[l] || cells(L) <= TAIL_CELLS || (g->tail_level >= 0 && (int)l >= g->tail_level)) continue;
      auto pick = [](int n, int first) {
         for (int t = first; t >= 2; t /= 2)
            if (n % t == 0) return t;
         return 0;
      };
      TileShape T;
      T.t[0] = pick(L.n[0], ndim == 3 ? 32 : 64);
      T.t[1] = pick(L.n[1], ndim == 3 ? 8 : 16);
      T.t[2] = ndim == 3 ? pick(L.n[2], 8) : 1;
      if (T.t[0] < 8 || T.t[1] < 2 || T.t[2] < 1) continue;  // slivers: the halo would dominate
      if (cudaMalloc(&g->alt_u[l], sizeof(double) * cells(L) * ncomp) != cudaSuccess) {
         g->alt_u[l] = nullptr;
         cudaGetLastError();
         continue;
      }
      g->alt_owned[l] = g->alt_u[l];
      g->tile[l] = T;
   }
   *out = g;
   return AMPE_OK;
}

int ampe_mg_create(int ndim, const int* n, const double* dx, int with_column_scale, ampe_mg** out)
{
   return ampe_mg_create_multi(ndim, n, dx, with_column_scale, 1, out);
}

int ampe_mg_destroy(ampe_mg* g)
{
   if (!g) return AMPE_OK;
   if (g->graph_exec) cudaGraphExecDestroy(g->graph_exec);
   for (double* b : g->blocks) cudaFree(b);
   for (double* b : g->alt_owned) cudaFree(b);
   for (double* b : g->own_c) cudaFree(b);
   for (double* b : g->own_m) cudaFree(b);
   for (int a = 0; a < 3; a++)
      for (double* b : g->own_d[a]) cudaFree(b);
   delete g;
   return AMPE_OK;
}

int ampe_mg_num_levels(const ampe_mg* g) { return g ? (int)g->levels.size() : 0; }
int ampe_mg_num_components(const ampe_mg* g) { return g ? g->ncomp : 0; }
int ampe_mg_last_launch_count(const ampe_mg* g) { return g ? g->launches : 0; }

int ampe_mg_set_zero_slope(ampe_mg* g, const int* zero_slope)
{
   if (!g || !zero_slope) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_zero_slope: NULL argument");
   for (int d = 0; d < 3; d++) g->zero_slope[d] = (d < g->ndim && zero_slope[d]) ? 1 : 0;
   for (Level& L : g->levels)
      for (int d = 0; d < 3; d++) L.clamp[d] = g->zero_slope[d];
   g->coefficients_set = false;  // the next set_* call rebuilds the face coefficients with the boundary faces zeroed
   if (g->graph_exec) cudaGraphExecDestroy(g->graph_exec), g->graph_exec = nullptr;
   return AMPE_OK;
}

int ampe_mg_set_sweeps(ampe_mg* g, int pre, int post, int coarse)
{
   if (!g || pre < 0 || post < 0 || coarse < 1 || pre + post < 1)
      return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_sweeps: bad argument");
   g->pre = pre, g->post = post, g->coarse = coarse;
   return AMPE_OK;
}

int ampe_mg_level_extents(const ampe_mg* g, int level, int* n_out)
{
   if (!g || !n_out || level < 0 || level >= (int)g->levels.size())
      return ampe_set_err(AMPE_EINVAL, "ampe_mg_level_extents: bad argument");
   for (int d = 0; d < 3; d++) n_out[d] = g->levels[level].n[d];
   return AMPE_OK;
}

int ampe_mg_copy_level(ampe_mg* g, int level, int which, double* out, void* stream)
{
   if (!g || !out || level < 0 || level >= (int)g->levels.size() || which < 0 || which > 5)
      return ampe_set_err(AMPE_EINVAL, "ampe_mg_copy_level: bad argument");
   const Level& L = g->levels[level];
   const double* src = which == 0 ? L.c : which == 1 ? L.m : which == 2 ? L.s : L.d[which - 3];
   if (!src) {
      // a constant of the block: fill the caller's array with it
      if (which == 2) return ampe_set_err(AMPE_EINVAL, "ampe_mg_copy_level: this operator has no column scale");
      if (which >= 3 && which - 3 >= g->ndim) return ampe_set_err(AMPE_EINVAL, "ampe_mg_copy_level: no such direction");
      const double v = which == 0 ? L.c_const : which == 1 ? L.m_const : L.d_const[which - 3];
      std::vector<double> h((size_t)cells(L), v);
      CUDA_OKM(cudaMemcpy(out, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
      return AMPE_OK;
   }
   CUDA_OKM(cudaMemcpyAsync(out, src, sizeof(double) * cells(L), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
   return AMPE_OK;
}

int ampe_mg_set_elliptic(ampe_mg* g, const double* m, int ngm, double m_const, const double* c, int ngc,
                         double c_const, const double* const* d, const double* const* d2, int ngd,
                         double d_scale, double d_const, void* stream)
{
   if (!g) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_elliptic: NULL solver");
   if (g->with_s) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_elliptic: solver was created for the quaternion block");
   if (d2 && !d) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_elliptic: second diffusion array without the first");
   P3 p = {{nullptr, nullptr, nullptr}}, p2 = {{nullptr, nullptr, nullptr}};
   for (int a = 0; a < g->ndim; a++) {
      if (d) {
         if (!d[a]) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_elliptic: NULL diffusion side array");
         p.a[a] = d[a];
      }
      if (d2) {
         if (!d2[a]) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_elliptic: NULL diffusion side array");
         p2.a[a] = d2[a];
      }
   }
   cudaStream_t st = (cudaStream_t)stream;
   g->launches = 0;
   // a zero-slope boundary needs the face coefficients as arrays (their wrap faces are zeroed), also where D is
   // a constant of the block
   const bool bc = g->zero_slope[0] || g->zero_slope[1] || g->zero_slope[2];
   int rc = configure(g, c != nullptr, c_const, m != nullptr, m_const, d != nullptr || bc, d_const);
   if (rc) return rc;
   if (m || c || d || bc) {
      const Level& L = g->levels[0];
      if (m || c || d) {
         mg_set_elliptic_kernel<<<grid_for(cells(L)), MT, 0, st>>>(L, m, ngm, c, ngc, p, p2, d2 ? 1 : 0, ngd, d_scale,
                                                                   g->inv_h2[0], g->inv_h2[1], g->inv_h2[2]);
         g->launches += 1;
      }
      if (bc) {
         mg_boundary_faces_kernel<<<grid_for(cells(L)), MT, 0, st>>>(L, d == nullptr ? 1 : 0);
         g->launches += 1;
      }
      return build_coarse(g, st);
   }
   g->coefficients_set = true;  // every coefficient is a constant: nothing to compute on the device
   return AMPE_OK;
}

int ampe_mg_set_quat(ampe_mg* g, double gamma, const double* mobility, int ngm, const double* const* face_coef,
                     int ngfc, void* stream)
{
   if (!g || !mobility || !face_coef) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_quat: NULL argument");
   if (!g->with_s) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_quat: solver was created without the column scale");
   P3 p = {{nullptr, nullptr, nullptr}};
   for (int a = 0; a < g->ndim; a++) {
      if (!face_coef[a]) return ampe_set_err(AMPE_EINVAL, "ampe_mg_set_quat: NULL face coefficient array");
      p.a[a] = face_coef[a];
   }
   cudaStream_t st = (cudaStream_t)stream;
   g->launches = 0;
   int rc = configure(g, false, 1.0, true, 0.0, true, 0.0);
   if (rc) return rc;
   const Level& L = g->levels[0];
   mg_set_quat_kernel<<<grid_for(cells(L)), MT, 0, st>>>(L, gamma, mobility, ngm, p, ngfc, g->inv_h2[0],
                                                         g->inv_h2[1], g->inv_h2[2]);
   g->launches += 1;
   if (g->zero_slope[0] || g->zero_slope[1] || g->zero_slope[2]) {
      mg_boundary_faces_kernel<<<grid_for(cells(L)), MT, 0, st>>>(L, 0);
      g->launches += 1;
   }
   return build_coarse(g, st);
}

int ampe_mg_solve(ampe_mg* g, const double* rhs, double* soln, int ncycles, int symmetrized, void* stream)
{
   if (!g || !rhs || !soln || ncycles < 1) return ampe_set_err(AMPE_EINVAL, "ampe_mg_solve: bad argument");
   if (!g->coefficients_set) return ampe_set_err(AMPE_EINVAL, "ampe_mg_solve: operator coefficients not set");
   cudaStream_t st = (cudaStream_t)stream;
   const Level& L = g->levels[0];
   auto issue = [&](cudaStream_t s) {
      g->launches = 0;
      mg_load_kernel<<<grid_for(cells(L)), MT, 0, s>>>(L, rhs, symmetrized);
      for (int c = 0; c < ncycles; c++) vcycle(g, s);
      mg_store_kernel<<<grid_for(cells(L)), MT, 0, s>>>(L, soln, symmetrized);
      g->launches += 2;
   };
   if (g->use_graph && st != nullptr) {
      // re-captured when the vectors, the cycle count or the coefficients (ampe_mg_set_*) change
      const bool hit = g->graph_exec && g->graph_rhs == rhs && g->graph_soln == soln &&
                       g->graph_cycles == ncycles && g->graph_symm == symmetrized;
      if (!hit) {
         if (g->graph_exec) cudaGraphExecDestroy(g->graph_exec), g->graph_exec = nullptr;
         cudaGraph_t graph = nullptr;
         CUDA_OKM(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
         issue(st);
         CUDA_OKM(cudaStreamEndCapture(st, &graph));
         CUDA_OKM(cudaGraphInstantiate(&g->graph_exec, graph, 0));
         cudaGraphDestroy(graph);
         g->graph_rhs = rhs, g->graph_soln = soln, g->graph_cycles = ncycles, g->graph_symm = symmetrized;
         g->graph_launches = g->launches;
      }
      CUDA_OKM(cudaGraphLaunch(g->graph_exec, st));
      g->launches = g->graph_launches;
      return AMPE_OK;
   }
   issue(st);
   CUDA_OKM(cudaGetLastError());
   return AMPE_OK;
}

int ampe_mg_apply(ampe_mg* g, const double* u, double* out, void* stream)
{
   if (!g || !u || !out || u == out) return ampe_set_err(AMPE_EINVAL, "ampe_mg_apply: bad argument");
   if (!g->coefficients_set) return ampe_set_err(AMPE_EINVAL, "ampe_mg_apply: operator coefficients not set");
   const Level& L = g->levels[0];
   mg_apply_kernel<<<grid_for(cells(L)), MT, 0, (cudaStream_t)stream>>>(L, u, out);
   g->launches = 1;
   CUDA_OKM(cudaGetLastError());
   return AMPE_OK;
}

int ampe_k_phasefacops_setc(int ndim, const int* ifirst, const int* ilast, const double* phi, int ngphi,
                            const double* m, int ngm, double gamma, double phi_well_scale,
                            const char* phi_well_func_type, double* c, int ngc, void* stream)
{
   if (!ifirst || !ilast || !phi || !m || !c || !phi_well_func_type || (ndim != 2 && ndim != 3))
      return ampe_set_err(AMPE_EINVAL, "ampe_k_phasefacops_setc: bad argument");
   const char t = phi_well_func_type[0];
   if (t != 'd' && t != 's')
      return ampe_set_err(AMPE_EINVAL, "ampe_k_phasefacops_setc: well type must be 'd' or 's' (functions.f)");
   Level L;
   L.ndim = ndim;
   for (int d = 0; d < 3; d++) {
      if (d < ndim && ifirst[d] != 0)
         return ampe_set_err(AMPE_EINVAL, "ampe_k_phasefacops_setc: the level box starts at 0");
      L.n[d] = d < ndim ? ilast[d] - ifirst[d] + 1 : 1;
   }
   L.c = L.m = L.s = L.u = L.f = L.r = nullptr;
   L.d[0] = L.d[1] = L.d[2] = nullptr;
   L.nc = 1, L.cs = 0;
   L.clamp[0] = L.clamp[1] = L.clamp[2] = 0;
   phasefacops_setc_kernel<<<grid_for(cells(L)), MT, 0, (cudaStream_t)stream>>>(L, phi, ngphi, m, ngm, gamma,
                                                                              phi_well_scale, t, c, ngc);
   CUDA_OKM(cudaGetLastError());
   return AMPE_OK;
}

}  // extern "C"
