// Scalar energy diagnostics on the staged tile.
//
// QuatModel::evaluateEnergy (QuatModel.cc:4888-4976) -> TwoPhasesEnergyEvaluationStrategy
// -> quatenergy / phi_interface_energy / interface_anisotropic_energy / bulkenergy
// ({2d,3d}/quatenergy.m4).  The reference materialises the side gradients of q first
// (computeDiffs + computeGradSide); here every cell recomputes |grad q|^2 on its 2*NDIM faces
// from the staged tile (rhs_math.cuh face_grad2, symmetric variant included), so the only
// traffic is one read of the state.  Per-block partial sums go to a buffer that a second
// one-block kernel adds in a fixed order: the result is deterministic.
//
// The Cahn-Hilliard model has no energy evaluator in the reference; PFHub benchmark 1a defines
// F = sum [ w (c-ca)^2 (cb-c)^2 + kappa/2 |grad c|^2 ] dV (ch_energy_kernel, rhs_common.cuh side).
#pragma once
#include "rhs_tile.cuh"

namespace ampe {

#define AMPE_NENERGY 6  // total, phi interface, orientational, q interface, well, bulk free

template <int NT>
AMPE_DEV void block_reduce_store(double* v, double* out)
{
   __shared__ double red[AMPE_NENERGY][NT / 32];
   const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
   for (int n = 0; n < AMPE_NENERGY; n++) {
      double x = v[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) red[n][warp] = x;
   }
   __syncthreads();
   if (threadIdx.x < AMPE_NENERGY) {
      double x = 0.0;
      for (int w = 0; w < NT / 32; w++) x += red[threadIdx.x][w];
      out[threadIdx.x] = x;
   }
}

template <class TT>
__global__ void __launch_bounds__(TT::NT) energy_tile_kernel(const __grid_constant__ FusedArgs A)
{
   using R = Rhs3<TT>;
   using SEL = typename TT::SEL;
   constexpr int ND = TT::ND, Q = TT::Q, CONC = TT::CONC, S = TT::S, NT = TT::NT, NW = TT::NW;
   constexpr int TX = TT::TX, TY = TT::TY, TZ = TT::TZ, CPT = TT::CPT;
   constexpr bool SYMM = TT::SYMM, WT = TT::WT;
   const Params& p = A.p;
   extern __shared__ double smem[];
   double* s = smem;
   int* s_iq = reinterpret_cast<int*>(smem + TT::O_END);
   __shared__ double s_qr[SYMM ? 48 : 1][4];
   __shared__ int s_conj[SYMM ? 48 : 1];
   if (SYMM && Q == 4) {
      for (int t = threadIdx.x; t < 48 * 4; t += NT) s_qr[t / 4][t % 4] = A.qr[t];
      for (int t = threadIdx.x; t < 48; t += NT) s_conj[t] = A.conj[t];
   }
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ox = blockIdx.x * TX;
   const int oy = blockIdx.y * TY + ((ND == 2) ? A.s_begin : 0);
   const int oz = (ND == 3) ? (blockIdx.z * TZ + A.s_begin) : 0;
   const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
   stage_tile<TT>(A, s, s_iq, ox, oy, oz);
   __syncthreads();

   constexpr int RSTEP_J = (NW < TY) ? NW : 0;
   constexpr int RSTEP_K = (NW < TY) ? 0 : NW / TY;
   const int lj0 = warp % TY, lk0 = warp / TY;
   constexpr int CSTEP = RSTEP_J * TT::SX + RSTEP_K * TT::SX * TT::SY;
   constexpr ZOff ZT = {-TT::SX * TT::SY, TT::SX * TT::SY};

   double weight = p.h[0] * p.h[1];
   if (ND == 3) weight = weight * p.h[2];
   const double floor2 = (p.grad_floor_type == 's') ? p.floor2 : 0.0;
   double acc[AMPE_NENERGY];
#pragma unroll
   for (int n = 0; n < AMPE_NENERGY; n++) acc[n] = 0.0;

   int c = TT::sidx(lane, lj0, lk0);
   const int gi = ox + lane;
   int gj = oy + lj0, gk = oz + lk0;
   const double* sp = s + TT::O_PHI;
#pragma unroll 1
   for (int u = 0; u < CPT; u++, c += CSTEP, gj += RSTEP_J, gk += RSTEP_K) {
      bool ok = (gi < n0) && (gj < n1) && (gk < n2);
      if (ND == 2) ok = ok && (gj < A.s_end);
      if (ND == 3) ok = ok && (gk < A.s_end);
      if (!ok) continue;
      const double phi = sp[c];
      const double temp = WT ? s[TT::O_T + c] : p.T_uniform;
      double e_phi, e_or = 0.0, e_q = 0.0, e_free = 0.0;
      // ---- phase interface energy ----
      if (ND == 2 && Q > 0 && p.nu > 0.0) {
         // interface_anisotropic_energy (2d/quatenergy.m4:12-103)
         const double dphidx = (sp[c + 1] - sp[c - 1]) * p.p5inv[0];
         const double dphidy = (sp[c + TT::SX] - sp[c - TT::SX]) * p.p5inv[1];
         double q = s[TT::O_Q + c];
         q = fmin(1.0, fmax(-1.0, q));
         double sn, cs;
         aniso_trig_libm(dphidx, dphidy, q, p.knumber, Q, &sn, &cs);
         const double epstheta = p.epsilon_phase * (1.0 + p.nu * cs);
         e_phi = 0.5 * epstheta * epstheta * (dphidx * dphidx + dphidy * dphidy);
      } else {
         // phi_interface_energy (2d/quatenergy.m4:108-181)
         double d = p.dinv2[0] * (-sp[c + 1] + 2.0 * phi - sp[c - 1]) +
                    p.dinv2[1] * (-sp[c + TT::SX] + 2.0 * phi - sp[c - TT::SX]);
         if (ND == 3) d = d + p.dinv2[2] * (-sp[c + ZT.p] + 2.0 * phi - sp[c + ZT.m]);
         e_phi = (0.5 * p.epsilon_phase * p.epsilon_phase) * d * phi;
      }
      // ---- orientational and q interface energy: 2*ND faces of the cell ----
      if constexpr (Q > 0) if (AMPE_SEL(evolve_quat) && p.misorientation_factor > 0.0) {
         double sum_g2 = 0.0, eo = 0.0;
         auto face = [&](auto dir, bool upper) {
            constexpr int a = decltype(dir)::value;
            const int st = (a == 0) ? 1 : ((a == 1) ? TT::SX : TT::SX * TT::SY);
            const int cu = upper ? c + st : c, cm = cu - st;
            const double g2 = R::template face_grad2<a>(s, s_iq, s_qr, s_conj, p, cu, cm, ZT);
            sum_g2 += g2;
            const double aphi = average_func(sp[cm], sp[cu], p.avg_func);
            const double pphi = interp_func(aphi, p.orient_interp1);
            // 3d/quatenergy.m4:386-387: the upper y face takes the root before the floor is added
            const double o2 = (ND == 3 && a == 1 && upper) ? g2 : g2 + floor2;
            eo += sqrt(o2) * pphi;
         };
         using D0 = std::integral_constant<int, 0>;
         using D1 = std::integral_constant<int, 1>;
         face(D0(), false), face(D0(), true), face(D1(), false), face(D1(), true);
         if constexpr (ND == 3) {
            using D2 = std::integral_constant<int, 2>;
            face(D2(), false), face(D2(), true);
         }
         const double avgf = (ND == 2) ? 0.25 : (1.0 / 6.0);
         e_or = eo * temp * avgf * p.misorientation_factor;
         e_q = sum_g2 * avgf * p.epsilonq2_half * interp_func(phi, p.orient_interp2);
      }
      const double e_well = p.phi_well_scale * well_func(phi, 'd');
      // ---- bulkenergy with f_l(c_l), f_a(c_a) (computeFreeEnergyLiquid / SolidA) ----
      if constexpr (CONC != 0) {
         const double c_l = s[TT::O_CL + c], c_a = s[TT::O_CA + c];
         double f_l, f_a;
         if (p.free_energy == AMPE_FE_CALPHAD) {
            f_l = calphad_f(p.ct, c_l, 0);
            f_a = calphad_f(p.ct, c_a, 1);
         } else {
            f_l = p.quad_A[0] * (c_l - p.quad_ceq[0]) * (c_l - p.quad_ceq[0]);
            f_a = p.quad_A[1] * (c_a - p.quad_ceq[1]) * (c_a - p.quad_ceq[1]);
         }
         f_l *= p.inv_vm_l;
         f_a *= p.inv_vm_a;
         const double h = interp_func(phi, p.energy_interp);
         e_free = (1.0 - h) * f_l + h * f_a;
      }
      acc[1] += e_phi * weight;
      acc[2] += e_or * weight;
      acc[3] += e_q * weight;
      acc[4] += e_well * weight;
      acc[5] += e_free * weight;
   }
   acc[0] = acc[1] + acc[2] + acc[3] + acc[4] + acc[5];
   const long long blk = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
   block_reduce_store<NT>(acc, A.energy_partials + blk * AMPE_NENERGY);
}

}  // namespace ampe
