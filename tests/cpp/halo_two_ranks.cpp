// Two slab ranks as two PROCESSES, no Python, no MPI, no NCCL: the C-ABI ghost-plane exchange (ampe_halo_*,
// ampe_rhs_eval_slab, ampe_rhs_eval_slab_host; csrc/halo.cu) against the one-rank evaluation of the whole periodic
// domain, bit for bit.  The ranks are forked before CUDA is touched and ship the set-up handles over a socketpair --
// the role MPI_Sendrecv has in AMPE.  Works with one GPU (both ranks on device 0: CUDA IPC maps memory between
// processes on the same device) or two (rank r on device r).
//   g++ -std=c++17 -O1 -I include -I /usr/local/cuda/include tests/cpp/halo_two_ranks.cpp \
//       -L ampe_b200 -lampe_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/ampe_b200 -o /tmp/halo_two_ranks
// Models: Cahn-Hilliard (benchmarks/PFHub1a, ghost width 2) and Dendrite2D (phase + quaternion[2] + temperature,
// ghost width 1): parameters of ampe_b200/configs.py.
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ampe_b200.h"

#define CK(call)                                                                                  \
   do {                                                                                           \
      int rc_ = (call);                                                                           \
      if (rc_ != 0) {                                                                             \
         fprintf(stderr, "rank %d: %s failed (%d): %s\n", g_rank, #call, rc_, ampe_last_error()); \
         return 1;                                                                                \
      }                                                                                           \
   } while (0)
#define CU(call)                                                                              \
   do {                                                                                       \
      cudaError_t e_ = (call);                                                                \
      if (e_ != cudaSuccess) {                                                                \
         fprintf(stderr, "rank %d: %s: %s\n", g_rank, #call, cudaGetErrorString(e_));         \
         return 1;                                                                            \
      }                                                                                       \
   } while (0)

static int g_rank = 0;

static void base_config(ampe_rhs_config& c, int nx, int ny, double lo, double hi)
{
   memset(&c, 0, sizeof(c));
   c.ndim = 2;
   c.n[0] = nx, c.n[1] = ny, c.n[2] = 1;
   c.dx[0] = (hi - lo) / nx, c.dx[1] = (hi - lo) / nx, c.dx[2] = 1.0;  // square cells (ny is the slab count)
   c.lag_quat_sidegrad = 1;
   c.quat_grad_modulus_from_cells = 1;
   c.energy_interp = 'p', c.conc_interp = 'p', c.diffusion_interp = 'l';
   c.orient_interp1 = 'q', c.orient_interp2 = 'c';
   c.avg_func = 'h', c.conc_avg_func = 'h', c.grad_floor_type = 'm', c.quat_mobility_func = 'p';
   c.knumber = 4;
   c.min_quat_mobility = 1.0e-6, c.quat_grad_floor = 1.0e-2, c.quat_mobility_alt_scale = 1.0;
   c.conc_mobility = 1.0, c.ch_mobility = 1.0;
   c.newton_max_its = 20, c.newton_tol = 1.0e-8, c.newton_alpha = 1.0;
   c.cp = 1.0, c.vm_liquid = c.vm_solid = 1.0e-6;
   c.nranks = 1, c.rank = 0;
}
static void pfhub1a(ampe_rhs_config& c, int nx, int ny)
{
   base_config(c, nx, ny, 0.0, 200.0);
   c.with_concentration = 1;
   c.conc_rhs_form = AMPE_CONC_CAHN_HILLIARD;
   c.T_uniform = 1000.0;
   c.ch_ca = 0.3, c.ch_cb = 0.7, c.ch_well_scale = 5.0, c.ch_kappa = 2.0;
   c.conc_mobility = 5.0;
}
static void dendrite2d(ampe_rhs_config& c, int nx, int ny)
{
   base_config(c, nx, ny, -4.5, 4.5);
   c.qlen = 2, c.with_phase = 1, c.with_unsteady_temperature = 1, c.evolve_quat = 1;
   c.phase_flux_type = AMPE_FLUX_ANISOTROPIC, c.free_energy = AMPE_FE_BIASWELL;
   c.epsilon_anisotropy = 0.05, c.H_parameter = 0.001, c.epsilon_phase = 0.01, c.phi_mobility = 3333.3333;
   c.quat_mobility = 1.0, c.epsilon_q = 1.0e3, c.meltingT = 1.0, c.thermal_diffusivity = 1.0;
   c.latent_heat = 1.0, c.phi_well_scale = 0.015625, c.bias_well_alpha = 0.9, c.bias_well_gamma = 10.0;
}

struct Vec {  // one state / RHS vector on the device
   ampe_rhs_fields f;
   std::vector<double*> owned;
   int alloc(const ampe_rhs_config& c, size_t ncell)
   {
      memset(&f, 0, sizeof(f));
      auto mk = [&](double** p, size_t n) {
         if (cudaMalloc(p, n * sizeof(double)) != cudaSuccess) return 1;
         owned.push_back(*p);
         return 0;
      };
      int bad = 0;
      if (c.with_phase) bad |= mk(&f.phase, ncell);
      if (c.qlen > 0) bad |= mk(&f.quat, ncell * c.qlen);
      if (c.with_concentration) bad |= mk(&f.conc, ncell);
      if (c.with_unsteady_temperature) bad |= mk(&f.temperature, ncell);
      return bad;
   }
   ~Vec()
   {
      for (double* p : owned) cudaFree(p);
   }
};

// deterministic smooth fields on the GLOBAL grid (row j0 .. j0 + ny of a domain with nyg rows)
static void fill_host(const ampe_rhs_config& c, int nx, int ny, int j0, int nyg, std::vector<double>& phase,
                      std::vector<double>& quat, std::vector<double>& conc, std::vector<double>& temp)
{
   const size_t n = (size_t)nx * ny;
   phase.assign(c.with_phase ? n : 0, 0.0);
   quat.assign((size_t)c.qlen * n, 0.0);
   conc.assign(c.with_concentration ? n : 0, 0.0);
   temp.assign(c.with_unsteady_temperature ? n : 0, 0.0);
   for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx; i++) {
         const double x = (i + 0.5) / nx, y = (j0 + j + 0.5) / nyg;
         const size_t t = i + (size_t)nx * j;
         const double r = sqrt((x - 0.5) * (x - 0.5) + (y - 0.5) * (y - 0.5));
         if (c.with_phase) phase[t] = 0.5 * (1.0 - tanh((r - 0.2) / 0.05));
         if (c.qlen == 2) {
            const double a = 0.3 + 0.2 * cos(2 * M_PI * x) * cos(4 * M_PI * y);
            quat[t] = cos(a), quat[t + n] = sin(a);
         }
         if (c.with_concentration) conc[t] = 0.5 + 0.01 * (cos(6 * M_PI * x) * cos(4 * M_PI * y) + cos(2 * M_PI * (x - y)));
         if (c.with_unsteady_temperature) temp[t] = 0.5 + 0.5 * (c.with_phase ? phase[t] : 0.0);
      }
}
static int upload(const ampe_rhs_config& c, Vec& v, size_t n, const std::vector<double>& phase,
                  const std::vector<double>& quat, const std::vector<double>& conc, const std::vector<double>& temp)
{
   if (c.with_phase) CU(cudaMemcpy(v.f.phase, phase.data(), n * 8, cudaMemcpyHostToDevice));
   if (c.qlen > 0) CU(cudaMemcpy(v.f.quat, quat.data(), n * 8 * c.qlen, cudaMemcpyHostToDevice));
   if (c.with_concentration) CU(cudaMemcpy(v.f.conc, conc.data(), n * 8, cudaMemcpyHostToDevice));
   if (c.with_unsteady_temperature) CU(cudaMemcpy(v.f.temperature, temp.data(), n * 8, cudaMemcpyHostToDevice));
   return 0;
}
// compare the slab result with rows [j0, j0 + ns) of the whole-domain result, bit for bit
static int same_bits(const char* what, const double* slab_dev, const double* full_dev, size_t nslab, size_t off, int depth,
                     size_t nfull, bool slab_on_host)
{
   std::vector<double> a(nslab), b(nslab);
   int bad = 0;
   for (int m = 0; m < depth; m++) {
      if (slab_on_host)
         memcpy(a.data(), slab_dev + m * nslab, nslab * 8);
      else if (cudaMemcpy(a.data(), slab_dev + m * nslab, nslab * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
         return 1;
      if (cudaMemcpy(b.data(), full_dev + m * nfull + off, nslab * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
      if (memcmp(a.data(), b.data(), nslab * 8) != 0) {
         double worst = 0;
         for (size_t t = 0; t < nslab; t++) worst = fmax(worst, fabs(a[t] - b[t]));
         fprintf(stderr, "rank %d: %s component %d differs from the one-rank evaluation (max abs %.3e)\n", g_rank, what, m,
                 worst);
         bad = 1;
      }
   }
   return bad;
}

static int run_model(int rank, int sock, void (*build)(ampe_rhs_config&, int, int), const char* name, int nx, int ns)
{
   const int nyg = 2 * ns;
   ampe_rhs_config cf, cs;
   build(cf, nx, nyg);
   build(cs, nx, ns);
   cs.nranks = 2, cs.rank = rank;
   const size_t nfull = (size_t)nx * nyg, nslab = (size_t)nx * ns;
   std::vector<double> ph, q, cc, tt;
   // ---- reference: the whole domain on one rank
   ampe_rhs_ctx* full = nullptr;
   CK(ampe_rhs_create(&cf, &full));
   Vec yf, df;
   if (yf.alloc(cf, nfull) || df.alloc(cf, nfull)) return 1;
   fill_host(cf, nx, nyg, 0, nyg, ph, q, cc, tt);
   if (upload(cf, yf, nfull, ph, q, cc, tt)) return 1;
   CK(ampe_rhs_eval(full, 0.0, &yf.f, &df.f, 0, nullptr));
   CU(cudaDeviceSynchronize());
   // ---- my slab
   ampe_rhs_ctx* ctx = nullptr;
   CK(ampe_rhs_create(&cs, &ctx));
   ampe_halo* h = nullptr;
   CK(ampe_halo_create(ctx, rank, 2, &h));
   char mine[AMPE_HALO_HANDLE_BYTES], theirs[AMPE_HALO_HANDLE_BYTES];
   CK(ampe_halo_export(h, mine));
   if (write(sock, mine, sizeof(mine)) != (ssize_t)sizeof(mine) || read(sock, theirs, sizeof(theirs)) != (ssize_t)sizeof(theirs)) {
      fprintf(stderr, "rank %d: handle exchange failed\n", rank);
      return 1;
   }
   CK(ampe_halo_connect(h, theirs, theirs));  // two ranks: the lower and the upper neighbour are the same peer
   Vec ys, ds;
   if (ys.alloc(cs, nslab) || ds.alloc(cs, nslab)) return 1;
   fill_host(cs, nx, ns, rank * ns, nyg, ph, q, cc, tt);
   if (upload(cs, ys, nslab, ph, q, cc, tt)) return 1;
   int bad = 0;
   const size_t off = (size_t)rank * nslab;
   for (int rep = 0; rep < 4; rep++) {  // both buffer parities, twice
      CK(ampe_rhs_eval_slab(ctx, h, 0.0, &ys.f, &ds.f, 0, nullptr));
      CU(cudaDeviceSynchronize());
      if (cs.with_phase) bad |= same_bits("phase", ds.f.phase, df.f.phase, nslab, off, 1, nfull, false);
      if (cs.evolve_quat) bad |= same_bits("quat", ds.f.quat, df.f.quat, nslab, off, cs.qlen, nfull, false);
      if (cs.with_concentration) bad |= same_bits("conc", ds.f.conc, df.f.conc, nslab, off, 1, nfull, false);
      if (cs.with_unsteady_temperature) bad |= same_bits("temperature", ds.f.temperature, df.f.temperature, nslab, off, 1, nfull, false);
   }
   // ---- the host-buffer path of the same slab
   ampe_rhs_fields yh, dh;
   memset(&yh, 0, sizeof(yh));
   memset(&dh, 0, sizeof(dh));
   std::vector<double> o_ph(nslab, NAN), o_q(nslab * (cs.qlen > 0 ? cs.qlen : 1), NAN), o_c(nslab, NAN), o_t(nslab, NAN);
   if (cs.with_phase) yh.phase = ph.data(), dh.phase = o_ph.data();
   if (cs.qlen > 0) yh.quat = q.data(), dh.quat = o_q.data();
   if (cs.with_concentration) yh.conc = cc.data(), dh.conc = o_c.data();
   if (cs.with_unsteady_temperature) yh.temperature = tt.data(), dh.temperature = o_t.data();
   for (int rep = 0; rep < 2; rep++) {
      CK(ampe_rhs_eval_slab_host(ctx, h, 0.0, &yh, &dh, 0));
      if (cs.with_phase) bad |= same_bits("host phase", dh.phase, df.f.phase, nslab, off, 1, nfull, true);
      if (cs.evolve_quat) bad |= same_bits("host quat", dh.quat, df.f.quat, nslab, off, cs.qlen, nfull, true);
      if (cs.with_concentration) bad |= same_bits("host conc", dh.conc, df.f.conc, nslab, off, 1, nfull, true);
      if (cs.with_unsteady_temperature) bad |= same_bits("host temperature", dh.temperature, df.f.temperature, nslab, off, 1, nfull, true);
   }
   // both ranks are done with each other's buffers before either unmaps them
   char token = 1;
   if (write(sock, &token, 1) != 1 || read(sock, &token, 1) != 1) bad = 1;
   CK(ampe_halo_destroy(h));
   CK(ampe_rhs_destroy(ctx));
   CK(ampe_rhs_destroy(full));
   printf("rank %d: %s %dx%d per rank: %s\n", rank, name, nx, ns, bad ? "MISMATCH" : "slab == whole domain, bit for bit");
   fflush(stdout);
   return bad;
}

int main(int argc, char** argv)
{
   int sv[2];
   if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) {
      perror("socketpair");
      return 2;
   }
   const pid_t child = fork();  // before any CUDA call: each process gets its own context
   if (child < 0) {
      perror("fork");
      return 2;
   }
   const int rank = child == 0 ? 1 : 0;
   g_rank = rank;
   const int sock = sv[rank];
   close(sv[1 - rank]);
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
      fprintf(stderr, "rank %d: no CUDA device\n", rank);
      return 2;
   }
   if (argc > 1) ndev = atoi(argv[1]) < ndev ? atoi(argv[1]) : ndev;
   if (cudaSetDevice(rank % ndev) != cudaSuccess) return 2;
   int bad = 0;
   bad |= run_model(rank, sock, pfhub1a, "Cahn-Hilliard (ghost width 2)", 96, 24);
   bad |= run_model(rank, sock, dendrite2d, "Dendrite2D (phase + quaternion + temperature)", 128, 48);
   if (rank == 1) _exit(bad ? 1 : 0);
   int status = 0;
   waitpid(child, &status, 0);
   if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad = 1;
   printf("HALO TWO RANKS %s (%d device%s)\n", bad ? "FAILED" : "OK", ndev, ndev > 1 ? "s" : "");
   return bad;
}
