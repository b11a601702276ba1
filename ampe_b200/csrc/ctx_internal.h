// Private definition of the context behind the C ABI (shared by ctx.cu and vecops.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "rhs_common.cuh"
#include "tile_shape.h"

#define AMPE_MAX_HOST_CHUNKS 32

struct ampe_rhs_ctx {
   ampe_rhs_config cfg;
   ampe::Params p;
   int ns;            // planes along the slab axis
   long long plane;   // cells per plane
   long long ncell;
   int ng;
   double *cl = nullptr, *ca = nullptr, *cl_ref = nullptr, *ca_ref = nullptr;  // slab-ghosted
   double* df = nullptr;  // ghost-0 CALPHAD driving force (written by the KKS kernel)
   int* iq[3] = {nullptr, nullptr, nullptr};
   double* lagN[3] = {nullptr, nullptr, nullptr};
   double* lagD0[3] = {nullptr, nullptr, nullptr};
   double* lagD1[3] = {nullptr, nullptr, nullptr};
   int* nfail = nullptr;
   double* qr_dev = nullptr;  // 48x4 cubic symmetry rotations (setqr)
   int* conj_dev = nullptr;
   ampe_rhs_fields halo_lo, halo_hi;
   bool have_halo = false;
   // set by halo.cu around an evaluation whose boundary blocks wait for the ghost planes themselves
   const unsigned long long* wait_flag[2] = {nullptr, nullptr};
   unsigned long long wait_epoch = 0;
   bool have_ref = false;
   bool lag_valid = false;
   bool generic_only = false;  // AMPE_B200_GENERIC at create time
   bool split3d = false;       // AMPE_B200_SPLIT3D at create time (experiment, rhs_march.cuh)
   int launches = 0;
   // staging buffers for ampe_rhs_eval_host
   ampe_rhs_fields dev_y, dev_ydot;
   bool have_dev = false;
   // energy diagnostics / reductions (vecops.cu): per-block partial sums + device results
   double* partials = nullptr;
   long long partials_cap = 0;
   double* red_out = nullptr;
   // opt-in per-kernel timing of one evaluation (ampe_rhs_set_kernel_timing): events on the launching stream
   // before the KKS pre-pass, between it and the fused kernel, and after the fused kernel
   int* grain_label = nullptr;  // grains.cu scratch: labels, counts, flags, compacted grains
   int grain_cap = 0;
   bool time_kernels = false;
   cudaEvent_t ev_t[3] = {nullptr, nullptr, nullptr};
   cudaStream_t own_stream = nullptr, k_stream = nullptr, out_stream = nullptr;
   cudaEvent_t ev_in[AMPE_MAX_HOST_CHUNKS], ev_k[AMPE_MAX_HOST_CHUNKS];
};

int ampe_set_err(int code, const std::string& msg);
// per-block partial sums of the reductions: room for `nblocks` blocks of 8 doubles (vecops.cu)
int ampe_ensure_scratch(ampe_rhs_ctx* c, long long nblocks);
// one evaluation of the fused kernel family in "energy" mode (ctx.cu): fills c->partials
int ampe_launch_energy(ampe_rhs_ctx* c, const ampe_rhs_fields* y, cudaStream_t st, long long* nblocks);
