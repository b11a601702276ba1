// Grain diagnostics behind the C ABI: QuatModel::computeGrainDiagnostics (source/QuatModel.cc:2690-2705) ->
// Grains::findAndNumberGrains (source/Grains.cc:263-520) + Grains::computeGrainVolumes (Grains.cc:647-697).
//
// The reference gives every cell with phi >= phase_threshold its global cell index, then sweeps "take the lowest
// number among the face neighbours" until nothing changes (ghost numbers refilled per sweep: periodic directions wrap,
// physical boundaries stay -1), so a grain ends up numbered by the lowest cell index it contains; its volume is the sum
// of the control volumes of its cells.  The fixed point does not depend on the sweep order, so the device version is
// free to reach it differently: every pass takes the neighbours' minimum AND follows the label chain
// (label[label[..]]: a label is always the index of a cell of the same grain, and never larger than the cell's own),
// which converges in O(log diameter) passes instead of O(diameter) -- a 1024^3 grain would need thousands of sweeps.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "ctx_internal.h"

namespace {

struct GrainArgs {
   int n[3];
   int periodic[3];
   long long ncell;
   const double* phase;
   double threshold;
   int* label;
   int* changed;
};

__global__ void grain_init_kernel(GrainArgs A)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.ncell; i += (long long)gridDim.x * blockDim.x)
      A.label[i] = (A.phase[i] >= A.threshold) ? (int)i : -1;  // Grains.cc:353-358
}

__global__ void grain_sweep_kernel(GrainArgs A)
{
   const long long s1 = A.n[0], s2 = (long long)A.n[0] * A.n[1];
   bool any = false;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.ncell; i += (long long)gridDim.x * blockDim.x) {
      const int own = A.label[i];
      if (own < 0) continue;
      int m = own;
      const int x = (int)(i % s1), y = (int)((i / s1) % A.n[1]), z = (int)(i / s2);
      const int c[3] = {x, y, z};
      const long long st[3] = {1, s1, s2};
#pragma unroll
      for (int d = 0; d < 3; d++) {
         if (A.n[d] == 1) continue;
         // lower / upper face neighbour; across a periodic boundary the opposite cell, across a physical one nothing
         long long lo = i - st[d], hi = i + st[d];
         bool has_lo = true, has_hi = true;
         if (c[d] == 0) {
            lo = i + (long long)(A.n[d] - 1) * st[d];
            has_lo = A.periodic[d] != 0;
         }
         if (c[d] == A.n[d] - 1) {
            hi = i - (long long)(A.n[d] - 1) * st[d];
            has_hi = A.periodic[d] != 0;
         }
         if (has_lo) {
            const int v = A.label[lo];
            if (v >= 0 && v < m) m = v;
         }
         if (has_hi) {
            const int v = A.label[hi];
            if (v >= 0 && v < m) m = v;
         }
      }
      // follow the chain of labels (each one a cell of this grain with a label not larger than itself)
      for (int hop = 0; hop < 64; hop++) {
         const int v = A.label[m];
         if (v >= m) break;  // (v < 0 cannot happen: labelled cells never lose their label)
         m = v;
      }
      if (m < own) {
         A.label[i] = m;  // racing writers only ever lower a label; the fixed point is unique
         any = true;
      }
   }
   if (any) *A.changed = 1;
}

__global__ void grain_count_kernel(const int* label, long long ncell, int* count)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) {
      const int l = label[i];
      if (l >= 0) atomicAdd(&count[l], 1);
   }
}

__global__ void grain_collect_kernel(const int* label, const int* count, long long ncell, int max_grains, int* nfound,
                                     int* ids, int* cells)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) {
      if (label[i] == (int)i) {
         const int slot = atomicAdd(nfound, 1);
         if (slot < max_grains) ids[slot] = (int)i, cells[slot] = count[i];
      }
   }
}

}  // namespace

#define CUDA_OKG(x)                                                     \
   do {                                                                 \
      cudaError_t e_ = (x);                                             \
      if (e_ != cudaSuccess) return ampe_set_err(AMPE_ECUDA, cudaGetErrorString(e_)); \
   } while (0)

extern "C" int ampe_grain_volumes(ampe_rhs_ctx* c, const ampe_rhs_fields* y, double phase_threshold, int max_grains,
                                  int* ngrains, int* grain_ids, double* volumes, void* stream)
{
   if (!c || !y || !ngrains || !grain_ids || !volumes || max_grains < 1) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (!c->p.with_phase || !y->phase) return ampe_set_err(AMPE_EINVAL, "grain diagnostics need the phase field");
   if (!(phase_threshold > 0.0)) return ampe_set_err(AMPE_EINVAL, "grain diagnostics: phase_threshold > 0");  // Grains.cc:268
   if (c->cfg.nranks > 1) return ampe_set_err(AMPE_EINVAL, "grain diagnostics: single rank (this build)");
   if (c->ncell >= (1ll << 31)) return ampe_set_err(AMPE_EINVAL, "grain diagnostics: more than 2^31 cells");
   cudaStream_t st = (cudaStream_t)stream;
   const long long nc = c->ncell;
   // scratch: labels + per-label cell counts + {changed, found} + the compacted grains
   if (!c->grain_label || c->grain_cap < max_grains) {
      if (c->grain_label) cudaFree(c->grain_label);
      c->grain_label = nullptr;
      const size_t bytes = sizeof(int) * (size_t)(2 * nc + 2 + 2 * (long long)max_grains);
      CUDA_OKG(cudaMalloc(&c->grain_label, bytes));
      c->grain_cap = max_grains;
   }
   int* label = c->grain_label;
   int* count = label + nc;
   int* flags = count + nc;  // [0] changed, [1] grains found
   int* ids = flags + 2;
   int* cells = ids + max_grains;
   GrainArgs A;
   for (int d = 0; d < 3; d++) {
      A.n[d] = c->p.n[d];
      A.periodic[d] = c->p.clamp[d] ? 0 : 1;
   }
   A.ncell = nc;
   A.phase = y->phase;
   A.threshold = phase_threshold;
   A.label = label;
   A.changed = flags;
   const int blocks = (int)std::min<long long>((nc + 255) / 256, 148 * 16);
   grain_init_kernel<<<blocks, 256, 0, st>>>(A);
   // the reference bounds its sweeps by 4 x the widest extent (Grains.cc:314-324); chain following needs far fewer
   int width = 0;
   for (int d = 0; d < 3; d++) width = std::max(width, A.n[d]);
   const int max_passes = 4 * width + 8;
   int pass = 0, changed = 1;
   while (changed && pass < max_passes) {
      CUDA_OKG(cudaMemsetAsync(flags, 0, sizeof(int), st));
      for (int k = 0; k < 4; k++, pass++) grain_sweep_kernel<<<blocks, 256, 0, st>>>(A);  // read back every 4 passes
      CUDA_OKG(cudaMemcpyAsync(&changed, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_OKG(cudaStreamSynchronize(st));
   }
   if (changed) return ampe_set_err(AMPE_EINVAL, "grain diagnostics: numbering did not converge");
   CUDA_OKG(cudaMemsetAsync(count, 0, sizeof(int) * (size_t)(nc + 2), st));
   grain_count_kernel<<<blocks, 256, 0, st>>>(label, nc, count);
   grain_collect_kernel<<<blocks, 256, 0, st>>>(label, count, nc, max_grains, flags + 1, ids, cells);
   CUDA_OKG(cudaGetLastError());
   int found = 0;
   CUDA_OKG(cudaMemcpyAsync(&found, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
   CUDA_OKG(cudaStreamSynchronize(st));
   if (found > max_grains) {
      *ngrains = found;
      return ampe_set_err(AMPE_EINVAL, "grain diagnostics: more grains than max_grains (count returned in *ngrains)");
   }
   std::vector<int> hid(found), hcells(found);
   if (found > 0) {
      CUDA_OKG(cudaMemcpyAsync(hid.data(), ids, sizeof(int) * found, cudaMemcpyDeviceToHost, st));
      CUDA_OKG(cudaMemcpyAsync(hcells.data(), cells, sizeof(int) * found, cudaMemcpyDeviceToHost, st));
      CUDA_OKG(cudaStreamSynchronize(st));
   }
   // std::map order of the reference's printout: ascending grain number; volume = cells x control volume
   // (the reference adds the same control volume cell by cell: equal to 1e-13 relative for any grain size)
   std::vector<int> order(found);
   for (int i = 0; i < found; i++) order[i] = i;
   std::sort(order.begin(), order.end(), [&](int a, int b) { return hid[a] < hid[b]; });
   double dv = 1.0;
   for (int d = 0; d < c->cfg.ndim; d++) dv *= c->cfg.dx[d];
   for (int i = 0; i < found; i++) {
      grain_ids[i] = hid[order[i]];
      volumes[i] = (double)hcells[order[i]] * dv;
   }
   *ngrains = found;
   return AMPE_OK;
}

extern "C" int ampe_grain_numbers(ampe_rhs_ctx* c, int* grain_number)
{
   if (!c || !grain_number || !c->grain_label) return ampe_set_err(AMPE_EINVAL, "ampe_grain_volumes first");
   CUDA_OKG(cudaMemcpy(grain_number, c->grain_label, sizeof(int) * (size_t)c->ncell, cudaMemcpyDeviceToDevice));
   return AMPE_OK;
}
