// Slab ghost-plane exchange between GPUs behind the C ABI (include/ampe_b200.h, "slab ghost-plane exchange").
// Replaces, on the RHS path, what QuatIntegrator::fillScratch gets from SAMRAI's RefineSchedule::fillData over MPI
// (source/QuatIntegrator.cc:2873-2955): the neighbours' boundary planes of every state component.
//
// Design (B200 / NVLink 5): every rank owns ONE device allocation -- arrival flags + receive buffers, two epoch
// parities x {lower, upper ghost planes} x component x (nghosts planes) -- which its two neighbours map (CUDA IPC
// across processes, plain peer access inside one process).  Per evaluation:
//   halo_push_kernel   my lowest / highest planes of y -> the neighbours' buffers (16-byte NVLink stores from a
//                      grid-stride loop over all components: the planes of a slab are contiguous, nothing is
//                      packed), system-scope fence, last block writes the epoch into the neighbours' flags;
//   halo_wait_kernel   two threads spin on this rank's own flags (local HBM) until both neighbours' epochs arrived.
// Both are stream-ordered: no host synchronisation, no library collective, no packing kernel.  The fused kernels
// read the ghost planes from LOCAL memory afterwards (ampe_rhs_set_halo layout), so their inner loops never see
// NVLink latency.  Buffers are double-buffered by epoch parity: a neighbour can only push epoch e+2 after it has
// waited for my epoch e+1, which I push after my own boundary evaluation of epoch e -- the reuse of a parity is
// ordered without a second handshake.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ctx_internal.h"

using namespace ampe;

#define CUDA_OKH(call)                                                                      \
   do {                                                                                     \
      cudaError_t e_ = (call);                                                              \
      if (e_ != cudaSuccess)                                                                \
         return ampe_set_err(AMPE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
   } while (0)

namespace {

constexpr int MAX_SEG = 16;        // copies of one push: 2 sides x up to 8 components
constexpr int AUX_COMPS = 4;       // components of the auxiliary channel (reference concentrations, rotations)
constexpr size_t FLAG_STRIDE = 64; // bytes between flags
constexpr size_t FLAG_BYTES = 512;
constexpr unsigned long long WAIT_TIMEOUT_NS = 300ull * 1000ull * 1000ull * 1000ull;  // a rank may still be setting up

struct PushArgs {
   const char* src[MAX_SEG];
   char* dst[MAX_SEG];
   unsigned long long nbytes;  // per segment
   int nseg;
   int vec16;                  // every pointer and nbytes are multiples of 16
   unsigned long long* flag[2];
   unsigned long long value;
   unsigned int* counter;
};

__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ PushArgs A)
{
   const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
   if (A.vec16) {
      const size_t n = A.nbytes / 16;
      for (int s = 0; s < A.nseg; s++) {
         const uint4* src = reinterpret_cast<const uint4*>(A.src[s]);
         uint4* dst = reinterpret_cast<uint4*>(A.dst[s]);
         for (size_t i = tid; i < n; i += nth) dst[i] = src[i];
      }
   } else {
      const size_t n = A.nbytes / 4;
      for (int s = 0; s < A.nseg; s++) {
         const unsigned* src = reinterpret_cast<const unsigned*>(A.src[s]);
         unsigned* dst = reinterpret_cast<unsigned*>(A.dst[s]);
         for (size_t i = tid; i < n; i += nth) dst[i] = src[i];
      }
   }
   if (!A.flag[0] && !A.flag[1]) return;  // local copy: no arrival flag
   // the planes must be visible to the neighbour before the flag: every thread fences its own stores at system
   // scope, the last block to arrive at the counter publishes the epoch
   __threadfence_system();
   __syncthreads();
   if (threadIdx.x == 0) {
      const unsigned done = atomicAdd(A.counter, 1u);
      if (done == gridDim.x - 1) {
         atomicExch(A.counter, 0u);
         __threadfence_system();
         for (int f = 0; f < 2; f++)
            if (A.flag[f]) *reinterpret_cast<volatile unsigned long long*>(A.flag[f]) = A.value;
         __threadfence_system();
      }
   }
}

__global__ void halo_wait_kernel(const unsigned long long* f0, const unsigned long long* f1, unsigned long long epoch)
{
   const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(threadIdx.x == 0 ? f0 : f1);
   unsigned long long t0, t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
   while (*f < epoch) {
      __nanosleep(100);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > WAIT_TIMEOUT_NS) __trap();  // a neighbour died: fail loudly instead of hanging the GPU
   }
   __threadfence_system();
}

struct Handle {  // what ampe_halo_export writes (<= AMPE_HALO_HANDLE_BYTES)
   unsigned magic;
   int pid;
   int device;
   int ncomp;
   unsigned long long slot_bytes, region_bytes;
   unsigned long long addr;  // same-process peers use the pointer itself
   cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(Handle) <= AMPE_HALO_HANDLE_BYTES, "handle blob too small");
constexpr unsigned MAGIC = 0x414d5048u;  // "AMPH"

}  // namespace

struct ampe_halo {
   ampe_rhs_ctx* c = nullptr;
   int rank = 0, nranks = 1, dev = 0;
   int ncomp = 0;                 // state components exchanged per evaluation
   size_t slot_bytes = 0;         // nghosts planes of one component
   size_t off_chan[2] = {0, 0};   // byte offsets of the two channels' buffers in the region
   int comps[2] = {0, AUX_COMPS};
   size_t region_bytes = 0;
   char* region = nullptr;
   char* peer[2] = {nullptr, nullptr};  // mapped regions of the lower / upper neighbour
   bool opened[2] = {false, false};
   unsigned int* counter = nullptr;
   unsigned long long epoch_push[2] = {0, 0}, epoch_wait[2] = {0, 0};
   int pushed[2] = {0, 0};        // sides of the current epoch already pushed
   cudaStream_t comm = nullptr;
   cudaEvent_t ev_ready = nullptr, ev_arrived = nullptr;
   int launches = 0;
   bool connected = false;

   char* buf(char* base, int ch, int parity, int side, int comp) const
   {
      return base + off_chan[ch] + ((size_t)(parity * 2 + side) * comps[ch] + comp) * slot_bytes;
   }
   unsigned long long* flag(char* base, int ch, int side) const
   {
      return reinterpret_cast<unsigned long long*>(base + (size_t)(ch * 2 + side) * FLAG_STRIDE);
   }
};

namespace {

// component pointers of a state vector in the order of ampe_rhs_fields (depth components of quat unrolled)
int list_components(const ampe_rhs_ctx* c, const ampe_rhs_fields* y, const double** out)
{
   const Params& p = c->p;
   int n = 0;
   if (p.with_phase) out[n++] = y ? y->phase : nullptr;
   for (int m = 0; m < p.qlen; m++) out[n++] = (y && y->quat) ? y->quat + (long long)m * c->ncell : nullptr;
   if (p.with_conc) out[n++] = y ? y->conc : nullptr;
   if (p.with_T) out[n++] = y ? y->temperature : nullptr;
   return n;
}

// point the context's halo at the receive buffers of one parity (ampe_rhs_set_halo layout: component stride =
// nghosts planes)
void select_parity(ampe_halo* h, int parity)
{
   ampe_rhs_ctx* c = h->c;
   const Params& p = c->p;
   ampe_rhs_fields lo, hi;
   memset(&lo, 0, sizeof(lo));
   memset(&hi, 0, sizeof(hi));
   int m = 0;
   auto at = [&](int side, int comp) { return reinterpret_cast<double*>(h->buf(h->region, 0, parity, side, comp)); };
   if (p.with_phase) lo.phase = at(0, m), hi.phase = at(1, m), m++;
   if (p.qlen > 0) lo.quat = at(0, m), hi.quat = at(1, m), m += p.qlen;
   if (p.with_conc) lo.conc = at(0, m), hi.conc = at(1, m), m++;
   if (p.with_T) lo.temperature = at(0, m), hi.temperature = at(1, m), m++;
   c->halo_lo = lo;
   c->halo_hi = hi;
   c->have_halo = true;
}

int launch_push(ampe_halo* h, PushArgs& A, cudaStream_t st)
{
   uintptr_t bits = (uintptr_t)A.nbytes;
   for (int s = 0; s < A.nseg; s++) bits |= (uintptr_t)A.src[s] | (uintptr_t)A.dst[s];
   A.vec16 = (bits % 16 == 0) ? 1 : 0;
   A.counter = h->counter;
   const size_t units = (size_t)A.nseg * (A.nbytes / (A.vec16 ? 16 : 4));
   long long blocks = (long long)((units + 256 * 8 - 1) / (256 * 8));  // ~8 units per thread
   if (blocks < 1) blocks = 1;
   if (blocks > 592) blocks = 592;  // 4 per SM: enough outstanding stores for NVLink, leaves SMs to the interior
   halo_push_kernel<<<(int)blocks, 256, 0, st>>>(A);
   CUDA_OKH(cudaGetLastError());
   h->launches++;
   return AMPE_OK;
}

// one channel's push of nf fields: low planes (src_lo) to the lower neighbour's UPPER ghost slots, high planes
// (src_hi) to the upper neighbour's LOWER ghost slots
int push_channel(ampe_halo* h, int ch, int nf, const char* const* src_lo, const char* const* src_hi, size_t nbytes,
                 int sides, cudaStream_t st)
{
   if (!h->connected) return ampe_set_err(AMPE_EINVAL, "ampe_halo: not connected");
   if (nf > h->comps[ch] || 2 * nf > MAX_SEG) return ampe_set_err(AMPE_EINVAL, "ampe_halo: too many components");
   if (nbytes > h->slot_bytes) return ampe_set_err(AMPE_EINVAL, "ampe_halo: planes larger than the receive slots");
   sides &= 3 & ~h->pushed[ch];
   if (!sides) return AMPE_OK;
   const unsigned long long e = h->epoch_push[ch] + 1;
   const int parity = (int)(e & 1);
   PushArgs A;
   memset(&A, 0, sizeof(A));
   A.nbytes = nbytes;
   A.value = e;
   // zero-slope physical boundary along the slab axis (ghost width 1): the first rank's lower ghost plane is its own
   // lowest plane, the last rank's upper ghost plane its own highest one -- the ring is cut there and the plane goes
   // into this rank's own receive slot (same flags, same epochs: every rank still gets two arrivals per exchange)
   const bool cut = h->c->p.clamp[h->c->p.ndim - 1] != 0;
   const bool self_lo = cut && h->rank == 0, self_hi = cut && h->rank == h->nranks - 1;
   for (int f = 0; f < nf; f++) {
      if (sides & 1) {
         A.src[A.nseg] = src_lo[f];
         A.dst[A.nseg] = self_lo ? h->buf(h->region, ch, parity, 0, f) : h->buf(h->peer[0], ch, parity, 1, f);
         A.nseg++;
      }
      if (sides & 2) {
         A.src[A.nseg] = src_hi[f];
         A.dst[A.nseg] = self_hi ? h->buf(h->region, ch, parity, 1, f) : h->buf(h->peer[1], ch, parity, 0, f);
         A.nseg++;
      }
   }
   if (sides & 1) A.flag[0] = self_lo ? h->flag(h->region, ch, 0) : h->flag(h->peer[0], ch, 1);  // the lower neighbour's UPPER side
   if (sides & 2) A.flag[1] = self_hi ? h->flag(h->region, ch, 1) : h->flag(h->peer[1], ch, 0);
   int rc = launch_push(h, A, st);
   if (rc) return rc;
   h->pushed[ch] |= sides;
   if (h->pushed[ch] == 3) {
      h->pushed[ch] = 0;
      h->epoch_push[ch] = e;
   }
   return AMPE_OK;
}

int wait_channel(ampe_halo* h, int ch, cudaStream_t st)
{
   if (!h->connected) return ampe_set_err(AMPE_EINVAL, "ampe_halo: not connected");
   const unsigned long long e = ++h->epoch_wait[ch];
   halo_wait_kernel<<<1, 2, 0, st>>>(h->flag(h->region, ch, 0), h->flag(h->region, ch, 1), e);
   CUDA_OKH(cudaGetLastError());
   h->launches++;
   return AMPE_OK;
}

// complete exchange of up to AUX_COMPS fields on the auxiliary channel: push, wait, copy the arrived planes out
int exchange_aux(ampe_halo* h, int nf, const char* const* src_lo, const char* const* src_hi, size_t nbytes,
                 char* const* dst_lo_ghost, char* const* dst_hi_ghost, cudaStream_t st)
{
   int rc = push_channel(h, 1, nf, src_lo, src_hi, nbytes, 3, st);
   if (rc) return rc;
   rc = wait_channel(h, 1, st);
   if (rc) return rc;
   const int parity = (int)(h->epoch_wait[1] & 1);
   PushArgs A;
   memset(&A, 0, sizeof(A));
   A.nbytes = nbytes;
   for (int f = 0; f < nf; f++) {
      A.src[A.nseg] = h->buf(h->region, 1, parity, 0, f);
      A.dst[A.nseg++] = dst_lo_ghost[f];
      A.src[A.nseg] = h->buf(h->region, 1, parity, 1, f);
      A.dst[A.nseg++] = dst_hi_ghost[f];
   }
   return launch_push(h, A, st);
}

}  // namespace

extern "C" {

int ampe_halo_create(ampe_rhs_ctx* c, int rank, int nranks, ampe_halo** out)
{
   if (!c || !out || nranks < 2 || rank < 0 || rank >= nranks)
      return ampe_set_err(AMPE_EINVAL, "ampe_halo_create: needs a context and 0 <= rank < nranks, nranks >= 2");
   ampe_halo* h = new ampe_halo;
   h->c = c;
   h->rank = rank;
   h->nranks = nranks;
   cudaGetDevice(&h->dev);
   const double* tmp[8];
   h->ncomp = list_components(c, nullptr, tmp);
   h->comps[0] = h->ncomp;
   // exactly nghosts planes: the fused kernels address the components of the ghost buffers with that stride
   h->slot_bytes = (size_t)c->ng * c->plane * sizeof(double);
   h->off_chan[0] = FLAG_BYTES;
   h->off_chan[1] = h->off_chan[0] + (size_t)4 * h->comps[0] * h->slot_bytes;
   h->region_bytes = h->off_chan[1] + (size_t)4 * h->comps[1] * h->slot_bytes;
   cudaError_t e = cudaMalloc(&h->region, h->region_bytes);
   if (e == cudaSuccess) e = cudaMemset(h->region, 0, FLAG_BYTES);
   if (e == cudaSuccess) e = cudaMalloc(&h->counter, sizeof(unsigned int));
   if (e == cudaSuccess) e = cudaMemset(h->counter, 0, sizeof(unsigned int));
   int lo_prio = 0, hi_prio = 0;
   if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
   if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->comm, cudaStreamNonBlocking, hi_prio);
   if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming);
   if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_arrived, cudaEventDisableTiming);
   if (e == cudaSuccess) e = cudaDeviceSynchronize();  // the flags are zero before anybody maps them
   if (e != cudaSuccess) {
      ampe_halo_destroy(h);
      return ampe_set_err(AMPE_ECUDA, std::string("ampe_halo_create: ") + cudaGetErrorString(e));
   }
   *out = h;
   return AMPE_OK;
}

int ampe_halo_export(ampe_halo* h, void* handle)
{
   if (!h || !handle) return ampe_set_err(AMPE_EINVAL, "ampe_halo_export: null argument");
   Handle H;
   memset(&H, 0, sizeof(H));
   H.magic = MAGIC;
   H.pid = (int)getpid();
   H.device = h->dev;
   H.ncomp = h->ncomp;
   H.slot_bytes = h->slot_bytes;
   H.region_bytes = h->region_bytes;
   H.addr = (unsigned long long)(uintptr_t)h->region;
   CUDA_OKH(cudaIpcGetMemHandle(&H.ipc, h->region));
   memset(handle, 0, AMPE_HALO_HANDLE_BYTES);
   memcpy(handle, &H, sizeof(H));
   return AMPE_OK;
}

int ampe_halo_connect(ampe_halo* h, const void* handle_prev, const void* handle_next)
{
   if (!h || !handle_prev || !handle_next) return ampe_set_err(AMPE_EINVAL, "ampe_halo_connect: null argument");
   const void* hs[2] = {handle_prev, handle_next};
   for (int s = 0; s < 2; s++) {
      Handle H;
      memcpy(&H, hs[s], sizeof(H));
      if (H.magic != MAGIC || H.ncomp != h->ncomp || H.slot_bytes != h->slot_bytes || H.region_bytes != h->region_bytes)
         return ampe_set_err(AMPE_EINVAL, "ampe_halo_connect: the neighbour's handle does not describe the same slab layout");
      if (s == 1 && memcmp(hs[0], hs[1], sizeof(Handle)) == 0) {  // two ranks: both neighbours are the same peer
         h->peer[1] = h->peer[0];
         break;
      }
      if (H.pid == (int)getpid()) {
         // same process (one process driving several GPUs, or the two-rank unit test): the pointer is valid as is
         if (H.device != h->dev) {
            int can = 0;
            CUDA_OKH(cudaDeviceCanAccessPeer(&can, h->dev, H.device));
            if (!can) return ampe_set_err(AMPE_EINVAL, "ampe_halo_connect: no peer access between the two devices");
            cudaError_t e = cudaDeviceEnablePeerAccess(H.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
               return ampe_set_err(AMPE_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            (void)cudaGetLastError();
         }
         h->peer[s] = reinterpret_cast<char*>((uintptr_t)H.addr);
      } else {
         void* p = nullptr;
         CUDA_OKH(cudaIpcOpenMemHandle(&p, H.ipc, cudaIpcMemLazyEnablePeerAccess));
         h->peer[s] = static_cast<char*>(p);
         h->opened[s] = true;
      }
   }
   h->connected = true;
   select_parity(h, 0);
   return AMPE_OK;
}

int ampe_halo_destroy(ampe_halo* h)
{
   if (!h) return AMPE_OK;
   cudaDeviceSynchronize();
   for (int s = 0; s < 2; s++)
      if (h->opened[s] && h->peer[s]) cudaIpcCloseMemHandle(h->peer[s]);
   if (h->c && h->connected) h->c->have_halo = false;
   cudaFree(h->region);
   cudaFree(h->counter);
   if (h->comm) cudaStreamDestroy(h->comm);
   if (h->ev_ready) cudaEventDestroy(h->ev_ready);
   if (h->ev_arrived) cudaEventDestroy(h->ev_arrived);
   (void)cudaGetLastError();
   delete h;
   return AMPE_OK;
}

int ampe_halo_push(ampe_halo* h, const ampe_rhs_fields* y, int sides, void* stream)
{
   if (!h || !y) return ampe_set_err(AMPE_EINVAL, "ampe_halo_push: null argument");
   const ampe_rhs_ctx* c = h->c;
   const double* comp[8];
   const int n = list_components(c, y, comp);
   const char *lo[8], *hi[8];
   for (int m = 0; m < n; m++) {
      if (!comp[m]) return ampe_set_err(AMPE_EINVAL, "ampe_halo_push: a state component is missing");
      lo[m] = reinterpret_cast<const char*>(comp[m]);
      hi[m] = reinterpret_cast<const char*>(comp[m] + (long long)(c->ns - c->ng) * c->plane);
   }
   return push_channel(h, 0, n, lo, hi, (size_t)c->ng * c->plane * sizeof(double), sides, (cudaStream_t)stream);
}

int ampe_halo_wait(ampe_halo* h, void* stream)
{
   if (!h) return ampe_set_err(AMPE_EINVAL, "ampe_halo_wait: null argument");
   int rc = wait_channel(h, 0, (cudaStream_t)stream);
   if (rc) return rc;
   // kernels launched from here on read the ghost planes of this epoch
   select_parity(h, (int)(h->epoch_wait[0] & 1));
   return AMPE_OK;
}

int ampe_rhs_eval_slab(ampe_rhs_ctx* c, ampe_halo* h, double time, const ampe_rhs_fields* y,
                       const ampe_rhs_fields* ydot, int fd_flag, void* stream)
{
   if (!c || !h || h->c != c) return ampe_set_err(AMPE_EINVAL, "ampe_rhs_eval_slab: context / halo mismatch");
   cudaStream_t st = (cudaStream_t)stream;
   h->launches = 0;
   // large ghost messages (3D slabs: tens of MB per face) travel on their own high-priority stream while the
   // interior planes are evaluated; small ones (2D: a few rows) cost less than the second stream's hand-offs
   const size_t msg = (size_t)h->ncomp * c->ng * c->plane * sizeof(double);
   const char* force = getenv("AMPE_B200_HALO_OVERLAP");
   bool overlap = msg >= ((size_t)1 << 20) && c->ns >= 4 * c->ng;
   if (force) overlap = (force[0] == '1') && c->ns >= 4 * c->ng;
   int rc;
   // Models without a KKS pre-pass (one fused kernel per evaluation): the push runs on the exchange stream and
   // the kernel's own boundary blocks wait for the arrival flags (rhs_common.cuh wait_ghost_planes) -- no wait
   // launch, the planes travel while every other block computes.  AMPE_B200_HALO_INKERNEL=0 switches it off.
   const Params& p = c->p;
   const bool one_kernel = !(p.conc_form == AMPE_CONC_KKS || p.conc_form == AMPE_CONC_EBS) &&
                           p.conc_form != AMPE_CONC_CAHN_HILLIARD;
   const char* ik = getenv("AMPE_B200_HALO_INKERNEL");
   if (one_kernel && !(ik && ik[0] == '0') && !force) {
      CUDA_OKH(cudaEventRecord(h->ev_ready, st));  // y is final once `st` gets here
      CUDA_OKH(cudaStreamWaitEvent(h->comm, h->ev_ready, 0));
      rc = ampe_halo_push(h, y, 3, h->comm);
      if (rc) return rc;
      const unsigned long long e = ++h->epoch_wait[0];
      select_parity(h, (int)(e & 1));
      c->wait_flag[0] = h->flag(h->region, 0, 0);
      c->wait_flag[1] = h->flag(h->region, 0, 1);
      c->wait_epoch = e;
      CUDA_OKH(cudaEventRecord(h->ev_arrived, h->comm));  // here: "my push has read y"
      rc = ampe_rhs_eval(c, time, y, ydot, fd_flag, st);
      c->wait_epoch = 0;
      if (rc) return rc;
      // whatever follows on `st` may overwrite y (a time step): not before the push has read its boundary planes
      CUDA_OKH(cudaStreamWaitEvent(st, h->ev_arrived, 0));
      h->launches += ampe_rhs_last_launch_count(c);
      return AMPE_OK;
   }
   if (overlap) {
      CUDA_OKH(cudaEventRecord(h->ev_ready, st));  // y is final once `st` gets here
      CUDA_OKH(cudaStreamWaitEvent(h->comm, h->ev_ready, 0));
      rc = ampe_halo_push(h, y, 3, h->comm);
      if (rc) return rc;
      rc = wait_channel(h, 0, h->comm);
      if (rc) return rc;
      CUDA_OKH(cudaEventRecord(h->ev_arrived, h->comm));
      rc = ampe_rhs_eval_interior(c, time, y, ydot, fd_flag, st);
      if (rc) return rc;
      int n = ampe_rhs_last_launch_count(c);
      CUDA_OKH(cudaStreamWaitEvent(st, h->ev_arrived, 0));
      select_parity(h, (int)(h->epoch_wait[0] & 1));
      rc = ampe_rhs_eval_boundary(c, time, y, ydot, fd_flag, st);
      if (rc) return rc;
      h->launches += n + (ampe_rhs_last_launch_count(c) - n);
   } else {
      rc = ampe_halo_push(h, y, 3, st);
      if (rc) return rc;
      rc = ampe_halo_wait(h, st);
      if (rc) return rc;
      rc = ampe_rhs_eval(c, time, y, ydot, fd_flag, st);
      if (rc) return rc;
      h->launches += ampe_rhs_last_launch_count(c);
   }
   return AMPE_OK;
}

int ampe_halo_last_launch_count(const ampe_halo* h) { return h ? h->launches : 0; }

int ampe_rhs_set_ref_concentrations_slab(ampe_rhs_ctx* c, ampe_halo* h, const double* cl_ref, const double* ca_ref,
                                         void* stream)
{
   if (!c || !h || h->c != c || !c->cl_ref) return ampe_set_err(AMPE_EINVAL, "context has no phase concentrations");
   if (!cl_ref || !ca_ref) return ampe_rhs_set_ref_concentrations(c, nullptr, nullptr, stream);
   cudaStream_t st = (cudaStream_t)stream;
   const long long pl = c->plane;
   const int ng = c->ng, ns = c->ns;
   const size_t nb = (size_t)ng * pl * sizeof(double);
   double* dst[2] = {c->cl_ref, c->ca_ref};
   const double* src[2] = {cl_ref, ca_ref};
   const char *lo[2], *hi[2];
   char *glo[2], *ghi[2];
   for (int f = 0; f < 2; f++) {
      CUDA_OKH(cudaMemcpyAsync(dst[f] + (long long)ng * pl, src[f], sizeof(double) * pl * ns, cudaMemcpyDeviceToDevice, st));
      lo[f] = reinterpret_cast<const char*>(src[f]);
      hi[f] = reinterpret_cast<const char*>(src[f] + (long long)(ns - ng) * pl);
      glo[f] = reinterpret_cast<char*>(dst[f]);
      ghi[f] = reinterpret_cast<char*>(dst[f] + (long long)(ng + ns) * pl);
   }
   int rc = exchange_aux(h, 2, lo, hi, nb, glo, ghi, st);
   if (rc) return rc;
   c->have_ref = true;
   return AMPE_OK;
}

int ampe_rhs_set_symmetry_rotations_slab(ampe_rhs_ctx* c, ampe_halo* h, const int* const* iqrot, void* stream)
{
   if (!c || !h || h->c != c || !c->p.symm) return ampe_set_err(AMPE_EINVAL, "context is not symmetry aware");
   cudaStream_t st = (cudaStream_t)stream;
   const long long pl = c->plane;
   const int ng = c->ng, ns = c->ns, nd = c->p.ndim;
   const size_t nb = (size_t)ng * pl * sizeof(int);
   const char *lo[3], *hi[3];
   char *glo[3], *ghi[3];
   for (int d = 0; d < nd; d++) {
      if (!iqrot || !iqrot[d]) return ampe_set_err(AMPE_EINVAL, "rotation array missing");
      CUDA_OKH(cudaMemcpyAsync(c->iq[d] + (long long)ng * pl, iqrot[d], sizeof(int) * pl * ns, cudaMemcpyDeviceToDevice, st));
      lo[d] = reinterpret_cast<const char*>(iqrot[d]);
      hi[d] = reinterpret_cast<const char*>(iqrot[d] + (long long)(ns - ng) * pl);
      glo[d] = reinterpret_cast<char*>(c->iq[d]);
      ghi[d] = reinterpret_cast<char*>(c->iq[d] + (long long)(ng + ns) * pl);
   }
   return exchange_aux(h, nd, lo, hi, nb, glo, ghi, st);
}

// QuatModel::computeSymmetryRotations on a slab rank: exchange the ghost planes of y, search the rotation of every
// lower face of this rank's cells (the faces of plane 0 along the slab axis see the lower neighbour's highest
// plane), then fetch the ghost planes of the indices from the neighbours
int ampe_rhs_compute_symmetry_rotations_slab(ampe_rhs_ctx* c, ampe_halo* h, const ampe_rhs_fields* y, void* stream)
{
   if (!c || !h || h->c != c || !y) return ampe_set_err(AMPE_EINVAL, "null argument");
   if (!c->p.symm) return ampe_set_err(AMPE_EINVAL, "context is not symmetry aware");
   cudaStream_t st = (cudaStream_t)stream;
   int rc = ampe_halo_push(h, y, 3, st);
   if (rc) return rc;
   rc = ampe_halo_wait(h, st);
   if (rc) return rc;
   rc = ampe_rhs_compute_symmetry_rotations(c, y, stream);
   if (rc) return rc;
   const long long pl = c->plane;
   const int ng = c->ng, ns = c->ns, nd = c->p.ndim;
   const size_t nb = (size_t)ng * pl * sizeof(int);
   const char *lo[3], *hi[3];
   char *glo[3], *ghi[3];
   for (int d = 0; d < nd; d++) {
      int* interior = c->iq[d] + (long long)ng * pl;
      lo[d] = reinterpret_cast<const char*>(interior);
      hi[d] = reinterpret_cast<const char*>(interior + (long long)(ns - ng) * pl);
      glo[d] = reinterpret_cast<char*>(c->iq[d]);
      ghi[d] = reinterpret_cast<char*>(c->iq[d] + (long long)(ng + ns) * pl);
   }
   return exchange_aux(h, nd, lo, hi, nb, glo, ghi, st);
}

}  // extern "C"
