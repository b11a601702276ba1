// Shared declarations of the fused RHS path: argument records of the kernels, the
// quaternion symmetry rotation, cp.async helpers, and the two small kernels that run
// beside the fused one (per-cell KKS solve, Cahn-Hilliard RHS).
#pragma once
#include "calphad.cuh"
#include "params.h"
#include "pointwise.cuh"
#include "tile_shape.h"

namespace ampe {

// caller-owned ghost-0 field + ghost planes along the slab axis (last axis)
struct Field {
   const double* base;
   const double* lo;  // ng planes below plane 0
   const double* hi;  // ng planes above plane ns-1
   long long comp;    // component stride in base
   long long hcomp;   // component stride in lo/hi
};

struct FusedArgs {
   Params p;
   Field phi, T, q, conc;
   // ctx-owned, slab-ghosted (ns+2 planes, pointer at plane -1): c_l, c_a
   const double* cl;
   const double* ca;
   const int* iq[3];  // slab-ghosted symmetry rotation indices (lower faces)
   const double* qr;  // 48x4 rotation table (setqr, quat.f:165-286), device global
   const int* conj;   // conjugate index table
   double* out_phi;
   double* out_q;
   double* out_c;
   double* out_T;
   // lagged face data (QuatIntegrator.cc:2804-2809, 3268-3269): 1/|grad q|_floor per
   // face and the composition face diffusivities, (ns+1) planes along the slab axis
   double* lagN[3];
   double* lagD0[3];
   double* lagD1[3];
   int use_lag;    // fd_flag != 0 && lag_quat_sidegrad
   int write_lag;  // refresh the lagged data
   int s_begin, s_end;  // slab-axis range of cells to compute [begin, end)
   int force_generic;   // host side only: use the runtime-selector instantiation
   int wrap_slab;       // 1: no halo buffers, ghost planes = opposite interior planes (one rank)
   const double* df;    // CALPHAD driving force (f_l-f_a)-mu(c_l-c_a) per cell from the KKS kernel
   double* energy_partials;  // non-null: energy diagnostics instead of the RHS (energy_tile.cuh)
   int split3d;              // host side only: AMPE_B200_SPLIT3D, two launches per 3D EBS evaluation
   // slab ranks (halo.cu): arrival flags of the lower / upper neighbour's ghost planes and the epoch to wait
   // for.  wait_epoch != 0: the blocks whose stage touches a ghost plane wait for the flag themselves, so that
   // the neighbours' pushes overlap the evaluation of every other block without a separate wait launch.
   const unsigned long long* wait_flag[2];
   unsigned long long wait_epoch;
};

// zero-slope physical boundary in direction d (-DAMPE_NO_CLAMP compiles the periodic-only kernels: A/B builds)
#ifdef AMPE_NO_CLAMP
#define AMPE_CLAMP(d) 0
#else
#define AMPE_CLAMP(d) (p.clamp[d])
#endif

// block-level wait for the ghost planes this block is about to stage (lo: below plane 0, hi: above plane ns-1)
AMPE_DEV void wait_ghost_planes(const FusedArgs& A, bool lo, bool hi)
{
#ifdef AMPE_NO_INKERNEL_WAIT
   return;
#endif
   if (A.wait_epoch == 0) return;  // uniform: single rank, or the exchange was waited for on the stream
   if (lo || hi) {
      if (threadIdx.x == 0) {
         unsigned long long t0, t;
         asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
         for (int side = 0; side < 2; side++) {
            if (!(side == 0 ? lo : hi)) continue;
            const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(A.wait_flag[side]);
            while (*f < A.wait_epoch) {
               __nanosleep(64);
               asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
               if (t - t0 > 300ull * 1000ull * 1000ull * 1000ull) __trap();  // a neighbour died: fail loudly
            }
         }
         __threadfence_system();
      }
      __syncthreads();
   }
}

// -DAMPE_SYMM_NOINLINE: the qlen-4 rotation as ONE out-of-line function (arguments and result in registers).
// The symmetry-aware tile kernel inlines ~40 rotations per face pair: 5700 SASS instructions (91 KB), and ncu
// shows "no instruction" as the top stall of its face phase (profiles/r02a_ncu_full_auni2d.txt) -- the kernel
// is bounded by instruction fetch, not by arithmetic.
struct Quat4 {
   double a, b, c, d;
};
static __device__ __noinline__ Quat4 symm_rotate4_call(double q0, double q1, double q2, double q3, int iq,
                                                       const double (*s_qr)[4], const int* s_conj)
{
   if (iq < 0) iq = s_conj[-iq - 1];
   Quat4 r;
   if (iq == 1) {
      r.a = q0, r.b = q1, r.c = q2, r.d = q3;
   } else {
      const double* b = s_qr[iq - 1];
      r.a = q0 * b[0] - q1 * b[1] - q2 * b[2] - q3 * b[3];
      r.b = q0 * b[1] + q1 * b[0] + q2 * b[3] - q3 * b[2];
      r.c = q0 * b[2] + q2 * b[0] + q3 * b[1] - q1 * b[3];
      r.d = q0 * b[3] + q3 * b[0] + q1 * b[2] - q2 * b[1];
   }
   return r;
}

template <int Q>
AMPE_DEV void symm_rotate(const double* q, int iq, double* qp, const double (*s_qr)[4],
                          const int* s_conj)
{
   if (Q == 4) {
#ifdef AMPE_SYMM_NOINLINE
      const Quat4 r = symm_rotate4_call(q[0], q[1], q[2], q[3], iq, s_qr, s_conj);
      qp[0] = r.a, qp[1] = r.b, qp[2] = r.c, qp[3] = r.d;
#else
      if (iq < 0) iq = s_conj[-iq - 1];
      if (iq == 1) {
#pragma unroll
         for (int m = 0; m < 4; m++) qp[m] = q[m];
      } else {
         quatmult4(q, s_qr[iq - 1], qp);
      }
#endif
   } else if (Q == 2) {
      // quatsymmrotate2 (quat.f:449-520): rotations (1,0),(0,1),(-1,0),(0,-1); conj 1,4,3,2
      if (iq < 0) iq = (iq == -2) ? 4 : ((iq == -4) ? 2 : -iq);
      const double r0 = (iq == 1) ? 1.0 : ((iq == 3) ? -1.0 : 0.0);
      const double r1 = (iq == 2) ? 1.0 : ((iq == 4) ? -1.0 : 0.0);
      if (iq == 1) {
         qp[0] = q[0];
         qp[1] = q[1];
      } else {
         qp[0] = q[0] * r0 - q[1] * r1;
         qp[1] = q[0] * r1 + q[1] * r0;
      }
   }
}

// libm evaluation of the 2D anisotropy angle functions exactly as written in
// anisotropic_gradient_flux (2d/quatrhs.m4:192-214); selected with AMPE_B200_LIBM_TRIG=1
static __device__ __noinline__ void aniso_trig_libm(double dphidx, double dphidy, double qa,
                                                    int knumber, int qlen, double* sn, double* cs)
{
   double theta;
   if (fabs(dphidx) > (double)1.e-12f)
      theta = atan(dphidy / dphidx);
   else
      theta = 0.5 * 3.141592653589793;  // 4.d0*atan(1.d0)
   const double ang = (qlen == 4) ? 2.0 * acos(qa) : acos(qa);
   sincos(knumber * (theta - ang), sn, cs);
}

AMPE_DEV void cp_async8(void* smem_dst, const void* gmem_src)
{
   const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
AMPE_DEV void cp_async4(void* smem_dst, const void* gmem_src)
{
   const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
AMPE_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- KKS pre-pass: (c_l, c_a) per cell on the slab + its ghost planes -----------------
// CALPHADequilibriumPhaseConcentrationsStrategy.cc:162-454 (Newton, warm start from *_ref)
// QuadraticEquilibriumPhaseConcentrationsStrategy.cc:42-144 (closed form, appendix.tex:462-490)
struct KksArgs {
   Params p;
   Field phi, conc;
   const double* cl_ref;  // slab-ghosted
   const double* ca_ref;
   double* cl;            // slab-ghosted outputs
   double* ca;
   double* df;            // ghost-0: CALPHAD driving force per interior cell (may be null)
   int* nfail;
   int s_begin, s_end;    // slab index range incl. ghosts: [-1, ns+1)
};

template <int ND>
__global__ void __launch_bounds__(256, AMPE_KKS_MINB) kks_kernel(const __grid_constant__ KksArgs A)
{
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;
   const long long total = plane * (A.s_end - A.s_begin);
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      const int sl = (int)(idx / plane) + A.s_begin;
      const long long inplane = idx % plane;
      double phi, conc;
      if (sl < 0) {
         phi = A.phi.lo[(long long)(sl + 1) * plane + inplane];
         conc = A.conc.lo[(long long)(sl + 1) * plane + inplane];
      } else if (sl >= ns) {
         phi = A.phi.hi[(long long)(sl - ns) * plane + inplane];
         conc = A.conc.hi[(long long)(sl - ns) * plane + inplane];
      } else {
         phi = A.phi.base[(long long)sl * plane + inplane];
         conc = A.conc.base[(long long)sl * plane + inplane];
      }
      const long long og = (long long)(sl + 1) * plane + inplane;
      const double hphi = interp_func(phi, p.conc_interp);
      double x0, x1;
      if (p.free_energy == AMPE_FE_CALPHAD) {
         x0 = A.cl_ref[og];
         x1 = A.ca_ref[og];
         KksFinal fin;
         const int st = kks_newton(p.ct, conc, hphi, x0, x1, p.newton_tol, p.newton_max_its,
                                   p.newton_alpha, fin);
         if (st < 0) atomicAdd(A.nfail, 1);
         if (A.df && sl >= 0 && sl < ns)
            A.df[(long long)sl * plane + inplane] =
                calphad_driving_force(p.ct, x0, x1, fin, p.inv_vm_l, p.inv_vm_a);
      } else {
         const double h = clamp01(hphi);
         x0 = (conc - h * (p.quad_ceq[1] - p.quad_rla * p.quad_ceq[0])) /
              ((1.0 - h) + h * p.quad_rla);
         x1 = (conc - (1.0 - h) * (p.quad_ceq[0] - p.quad_ral * p.quad_ceq[1])) /
              ((1.0 - h) * p.quad_ral + h);
      }
      A.cl[og] = x0;
      A.ca[og] = x1;
   }
}

// ---- Cahn-Hilliard composition RHS (config C1), ghost width 2 ---------------------------
// add_cahnhilliarddoublewell_flux (2d/concentrationrhs.m4:85-137) + computerhsconcentration
struct ChArgs {
   Params p;
   Field conc;  // lo/hi hold 2 planes each
   double* out_c;
   int s_begin, s_end;
};

template <int ND>
__global__ void __launch_bounds__(256) ch_kernel(const __grid_constant__ ChArgs A)
{
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;
   const long long total = plane * (A.s_end - A.s_begin);
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
   auto mu = [&](int i, int j, int k) -> double {
      const double c = at(i, j, k);
      double lap = p.ch_dinv2[0] * (-2.0 * c + at(i - 1, j, k) + at(i + 1, j, k)) +
                   p.ch_dinv2[1] * (-2.0 * c + at(i, j - 1, k) + at(i, j + 1, k));
      if (ND == 3) lap = lap + p.ch_dinv2[2] * (-2.0 * c + at(i, j, k - 1) + at(i, j, k + 1));
      return 2.0 * p.ch_well_scale * (c - p.ch_ca) * (p.ch_cb - c) * (p.ch_cb + p.ch_ca - 2.0 * c) -
             p.ch_kappa * lap;
   };
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      const long long cell = idx + (long long)A.s_begin * plane;
      const int i = (int)(cell % n0);
      const int j = (int)((cell / n0) % n1);
      const int k = (int)(cell / ((long long)n0 * n1));
      const double mc = mu(i, j, k);
      // scatter order of the reference: flux(i) = (0 - M/h mu(i-1)) + M/h mu(i)
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < ND; a++) {
         const double mm = mu(i - (a == 0), j - (a == 1), k - (a == 2));
         const double mp = mu(i + (a == 0), j + (a == 1), k + (a == 2));
         const double flo = (0.0 - p.ch_mdinv[a] * mm) + p.ch_mdinv[a] * mc;
         const double fup = (0.0 - p.ch_mdinv[a] * mc) + p.ch_mdinv[a] * mp;
         s = (a == 0) ? p.dinv[a] * (fup - flo) : s + p.dinv[a] * (fup - flo);
      }
      A.out_c[cell] = p.conc_mobility * s;
   }
}

}  // namespace ampe
