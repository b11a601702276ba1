// Per-cell arithmetic of the block preconditioners: one periodic uniform
// level of AMPE's cell-centred block operators and the geometric multigrid that inverts them.
//
// Operators (reference):
//   scalar blocks  A u = M div(D grad u) + C u       EllipticFACOps.h:35, 2d/ellipticfacops.m4:16-56
//                                                    (efo_compfluxvardc2d) and :346-393 (efo_compresvarsca2d)
//     phase        M = phase mobility, C = 1 + gamma M w g''(phi), D = -gamma eps^2
//                                                    PhaseFACOps.cc:33-51, :100-186
//     composition  M = conc mobility, C = 1, D = -gamma D_pfm         ConcFACOps.cc:19-50
//     temperature  M = 1, C = 1, D = -gamma kappa                     QuatIntegrator.cc:3340-3346
//   quaternion     A w = w + gamma sqrt(m) div(fc grad(sqrt(m) w)), one matrix for every component
//                                                    2d/quatlevelsolver.m4:9-118 (set_j_ij2d, set_stencil2d)
// Both are written   (A u)_i = c_i u_i + m_i sum_faces d_f (s_nb u_nb - s_i u_i),   d_f = D_f / h_f^2,
// with s = 1 for the scalar blocks and c = 1, m = gamma sqrt(mobility), s = sqrt(mobility) for the
// quaternion block.
//
// Smoother: the reference's red-black Gauss-Seidel update (efo_rbgswithfluxmaxvardcvarsf2d,
// 2d/ellipticfacops.m4:60-130): u_i += (f - A u)_i / diag_i.  The reference hands the single level
// to hypre PFMG (not in its tree); here the level is inverted by V-cycles over rediscretised
// coarse levels (coefficients averaged, residual averaged, cell-centred (bi/tri)linear prolongation).
//
// The functions are __host__ __device__ so that the very same arithmetic can be looped on the host
// by the
// test infrastructure outside this package -- the product only calls them from the mg.cu kernels.
#pragma once
#ifdef __CUDACC__
#define MG_HD __host__ __device__ __forceinline__
#else
#define MG_HD inline
#endif

namespace ampe_mg_cell {

// A coefficient that is a constant of the block (phase: M and D; temperature: M, C and D; composition:
// M and C) is not stored: its pointer is null and the value sits in the level descriptor, so the
// sweeps do not read arrays of identical numbers from HBM.
struct Level {
   int ndim;
   int n[3];      // cells per direction (n[2] = 1 in 2D); periodic
   double* c;     // cell: C, or nullptr (= c_const)
   double* m;     // cell: row multiplier, or nullptr (= m_const)
   double* s;     // cell: column multiplier, or nullptr (= 1)
   double* d[3];  // lower-face coefficient per direction, divided by h^2, or nullptr (= d_const[a])
   double* u;     // solution / correction
   double* f;     // right-hand side
   double* r;     // residual (and Jacobi work array)
   double c_const, m_const, d_const[3];
   // u, f, r hold nc components (the qlen components of the quaternion block share one matrix: one pass
   // updates all of them and reads the coefficients once); component m starts at m * cs
   int nc;
   long long cs;
   // zero-slope (homogeneous Neumann) physical boundary per direction: the coefficient of the wrap face (lower face
   // of the cells with index 0) is zero on every level, so that the periodic index arithmetic of the sweeps never
   // couples the two ends, and the prolongation does not interpolate across the boundary
   int clamp[3];
};

MG_HD double mg_c(const Level& L, long long o) { return L.c ? L.c[o] : L.c_const; }
MG_HD double mg_m(const Level& L, long long o) { return L.m ? L.m[o] : L.m_const; }
MG_HD double mg_d(const Level& L, int a, long long o) { return L.d[a] ? L.d[a][o] : L.d_const[a]; }

MG_HD long long mg_index(const Level& L, int i, int j, int k)
{
   return (long long)i + (long long)L.n[0] * ((long long)j + (long long)L.n[1] * (long long)k);
}
MG_HD int mg_up(int i, int n) { return i + 1 == n ? 0 : i + 1; }
MG_HD int mg_dn(int i, int n) { return i == 0 ? n - 1 : i - 1; }

// sum over the faces of cell (i,j,k) of d_f (s_nb u_nb - s_i u_i), and sum of d_f
MG_HD void mg_face_sums(const Level& L, const double* u, int i, int j, int k, double& flux, double& dsum)
{
   const long long o = mg_index(L, i, j, k);
   const double si = L.s ? L.s[o] : 1.0;
   const double ui = si * u[o];
   flux = 0.0;
   dsum = 0.0;
   {
      const long long ou = mg_index(L, mg_up(i, L.n[0]), j, k), od = mg_index(L, mg_dn(i, L.n[0]), j, k);
      const double du = mg_d(L, 0, ou), dd = mg_d(L, 0, o);
      const double uu = (L.s ? L.s[ou] : 1.0) * u[ou], ud = (L.s ? L.s[od] : 1.0) * u[od];
      flux += du * (uu - ui) - dd * (ui - ud);
      dsum += du + dd;
   }
   {
      const long long ou = mg_index(L, i, mg_up(j, L.n[1]), k), od = mg_index(L, i, mg_dn(j, L.n[1]), k);
      const double du = mg_d(L, 1, ou), dd = mg_d(L, 1, o);
      const double uu = (L.s ? L.s[ou] : 1.0) * u[ou], ud = (L.s ? L.s[od] : 1.0) * u[od];
      flux += du * (uu - ui) - dd * (ui - ud);
      dsum += du + dd;
   }
   if (L.ndim == 3) {
      const long long ou = mg_index(L, i, j, mg_up(k, L.n[2])), od = mg_index(L, i, j, mg_dn(k, L.n[2]));
      const double du = mg_d(L, 2, ou), dd = mg_d(L, 2, o);
      const double uu = (L.s ? L.s[ou] : 1.0) * u[ou], ud = (L.s ? L.s[od] : 1.0) * u[od];
      flux += du * (uu - ui) - dd * (ui - ud);
      dsum += du + dd;
   }
}

// (A u)_i
MG_HD double mg_apply_cell(const Level& L, const double* u, int i, int j, int k)
{
   double flux, dsum;
   mg_face_sums(L, u, i, j, k, flux, dsum);
   const long long o = mg_index(L, i, j, k);
   return mg_c(L, o) * u[o] + mg_m(L, o) * flux;
}

// r_i = f_i - (A u)_i   (efo_compresvarsca2d)
MG_HD void mg_residual_cell(const Level& L, int i, int j, int k)
{
   const long long o = mg_index(L, i, j, k);
   for (int m = 0; m < L.nc; m++) {
      const long long q = m * L.cs;
      L.r[o + q] = L.f[o + q] - mg_apply_cell(L, L.u + q, i, j, k);
   }
}

// Gauss-Seidel update of one cell (efo_rbgswithfluxmaxvardcvarsf2d): u += residual / diagonal
MG_HD void mg_smooth_cell(const Level& L, int i, int j, int k)
{
   const long long o = mg_index(L, i, j, k);
   const double si = L.s ? L.s[o] : 1.0;
   for (int m = 0; m < L.nc; m++) {
      double* u = L.u + m * L.cs;
      double flux, dsum;
      mg_face_sums(L, u, i, j, k, flux, dsum);
      const double residual = L.f[o + m * L.cs] - (mg_c(L, o) * u[o] + mg_m(L, o) * flux);
      const double diag = mg_c(L, o) - mg_m(L, o) * si * dsum;
      u[o] += residual / diag;
   }
}

// ---- one red-black sweep in ONE pass over a tile (fused smoother) ---------------------------------
// The two colour half-sweeps each move every array in full (a 32 B sector holds two cells of each
// colour).  Here a tile of TX x TY (x TZ) cells is staged with a halo of two into `tile` (shared
// memory on the device), the red cells of the tile grown by one are updated in place, then the black
// cells of the tile, and the tile is written to u_out: u_in is only read, so neighbouring tiles see old
// values and the result is the one of the two half-sweeps bit for bit (the halo's red updates are
// recomputed by the neighbour from the same data).  The caller swaps u_in / u_out afterwards.
// Written as phases of thread-strided loops separated by MG_TILE_SYNC so that the host can run the very
// same index arithmetic with one "thread" (tid 0 of 1).  Requires even extents (two-colouring).
struct TileShape {
   int t[3];  // tile extents (t[2] = 1 in 2D)
};
MG_HD int mg_wrap(int i, int n)
{
   i %= n;
   return i < 0 ? i + n : i;
}
// Gauss-Seidel value of cell (gi,gj,gk) (global, in range) from the staged tile: same operation order
// as mg_face_sums + mg_smooth_cell
MG_HD double mg_gs_from_tile(const Level& L, const double* f, const double* tile, int p0, int p1, int li, int lj, int lk,
                             int gi, int gj, int gk)
{
   // tile index of local (li,lj,lk) with the halo of two: (li+2) + p0 ((lj+2) + p1 (lk+hz))
   const int hz = L.ndim == 3 ? 2 : 0;
   const long long tc = (long long)(li + 2) + (long long)p0 * ((lj + 2) + (long long)p1 * (lk + hz));
   const long long o = mg_index(L, gi, gj, gk);
   const double si = L.s ? L.s[o] : 1.0;
   const double uc = tile[tc];
   const double ui = si * uc;
   double flux = 0.0, dsum = 0.0;
   {
      const long long ou = mg_index(L, mg_up(gi, L.n[0]), gj, gk), od = mg_index(L, mg_dn(gi, L.n[0]), gj, gk);
      const double du = mg_d(L, 0, ou), dd = mg_d(L, 0, o);
      const double uu = (L.s ? L.s[ou] : 1.0) * tile[tc + 1], ud = (L.s ? L.s[od] : 1.0) * tile[tc - 1];
      flux += du * (uu - ui) - dd * (ui - ud);
      dsum += du + dd;
   }
   {
      const long long ou = mg_index(L, gi, mg_up(gj, L.n[1]), gk), od = mg_index(L, gi, mg_dn(gj, L.n[1]), gk);
      const double du = mg_d(L, 1, ou), dd = mg_d(L, 1, o);
      const double uu = (L.s ? L.s[ou] : 1.0) * tile[tc + p0], ud = (L.s ? L.s[od] : 1.0) * tile[tc - p0];
      flux += du * (uu - ui) - dd * (ui - ud);
      dsum += du + dd;
   }
   if (L.ndim == 3) {
      const long long ou = mg_index(L, gi, gj, mg_up(gk, L.n[2])), od = mg_index(L, gi, gj, mg_dn(gk, L.n[2]));
      const double du = mg_d(L, 2, ou), dd = mg_d(L, 2, o);
      const long long pz = (long long)p0 * p1;
      const double uu = (L.s ? L.s[ou] : 1.0) * tile[tc + pz], ud = (L.s ? L.s[od] : 1.0) * tile[tc - pz];
      flux += du * (uu - ui) - dd * (ui - ud);
      dsum += du + dd;
   }
   const double residual = f[o] - (mg_c(L, o) * uc + mg_m(L, o) * flux);
   const double diag = mg_c(L, o) - mg_m(L, o) * si * dsum;
   return uc + residual / diag;
}

#ifdef __CUDA_ARCH__
#define MG_TILE_SYNC() __syncthreads()
#else
#define MG_TILE_SYNC() ((void)0)
#endif

// one tile with origin (o0,o1,o2) (multiples of the tile extents; extents of L are multiples of them too)
// (one component: u_in, u_out and f point at it; the caller loops over the components of the level)
MG_HD void mg_rb_tile_pass(const Level& L, const double* f, const double* u_in, double* u_out, double* tile, TileShape T,
                           int o0, int o1, int o2, int tid, int nthreads)
{
   const int hz = L.ndim == 3 ? 2 : 0;
   const int p0 = T.t[0] + 4, p1 = T.t[1] + 4, p2 = T.t[2] + 2 * hz;
   // phase 0: stage u_in with the halo of two (periodic images)
   for (int t = tid; t < p0 * p1 * p2; t += nthreads) {
      const int a = t % p0, b = (t / p0) % p1, c = t / (p0 * p1);
      const int gi = mg_wrap(o0 + a - 2, L.n[0]), gj = mg_wrap(o1 + b - 2, L.n[1]);
      const int gk = L.ndim == 3 ? mg_wrap(o2 + c - 2, L.n[2]) : 0;
      tile[t] = u_in[mg_index(L, gi, gj, gk)];
   }
   MG_TILE_SYNC();
   // phases 1, 2: red cells of the tile grown by one, then black cells of the tile, in place
   for (int colour = 0; colour < 2; colour++) {
      const int g = colour == 0 ? 1 : 0;  // growth of the region
      const int gz = L.ndim == 3 ? g : 0;
      const int e0 = T.t[0] + 2 * g, e1 = T.t[1] + 2 * g, e2 = T.t[2] + 2 * gz;
      const int h0 = e0 >> 1;  // e0 is even: every row of the region holds h0 cells of each colour
      for (int t = tid; t < h0 * e1 * e2; t += nthreads) {
         const int ii = t % h0, lj = (t / h0) % e1 - g, lk = t / (h0 * e1) - gz;
         // first cell of the colour in this row: (o0 + li + o1 + lj + o2 + lk) & 1 == colour, li = 2 ii + a - g
         const int a = (colour + o0 + g + o1 + lj + o2 + lk + 8) & 1;
         const int li = 2 * ii + a - g;
         const int gi = mg_wrap(o0 + li, L.n[0]), gj = mg_wrap(o1 + lj, L.n[1]);
         const int gk = L.ndim == 3 ? mg_wrap(o2 + lk, L.n[2]) : 0;
         const double v = mg_gs_from_tile(L, f, tile, p0, p1, li, lj, lk, gi, gj, gk);
         tile[(long long)(li + 2) + (long long)p0 * ((lj + 2) + (long long)p1 * (lk + hz))] = v;
      }
      MG_TILE_SYNC();
   }
   // phase 3: the tile goes to u_out
   for (int t = tid; t < T.t[0] * T.t[1] * T.t[2]; t += nthreads) {
      const int li = t % T.t[0], lj = (t / T.t[0]) % T.t[1], lk = t / (T.t[0] * T.t[1]);
      u_out[mg_index(L, o0 + li, o1 + lj, L.ndim == 3 ? o2 + lk : 0)] =
          tile[(long long)(li + 2) + (long long)p0 * ((lj + 2) + (long long)p1 * (lk + hz))];
   }
   MG_TILE_SYNC();  // the tile may be re-staged for the next component
}

// damped Jacobi for levels whose periodic wrap breaks the two-colouring (an odd extent):
// u += omega r / diagonal with r computed beforehand by mg_residual_cell
MG_HD void mg_jacobi_cell(const Level& L, double omega, int i, int j, int k)
{
   double dsum = 0.0;
   const long long o = mg_index(L, i, j, k);
   dsum += mg_d(L, 0, mg_index(L, mg_up(i, L.n[0]), j, k)) + mg_d(L, 0, o);
   dsum += mg_d(L, 1, mg_index(L, i, mg_up(j, L.n[1]), k)) + mg_d(L, 1, o);
   if (L.ndim == 3) dsum += mg_d(L, 2, mg_index(L, i, j, mg_up(k, L.n[2]))) + mg_d(L, 2, o);
   const double si = L.s ? L.s[o] : 1.0;
   const double diag = mg_c(L, o) - mg_m(L, o) * si * dsum;
   for (int m = 0; m < L.nc; m++) L.u[o + m * L.cs] += omega * L.r[o + m * L.cs] / diag;
}

// coarse cell (I,J,K): f_c = mean of the children's residuals, u_c = 0
MG_HD void mg_restrict_cell(const Level& F, const Level& C, int I, int J, int K)
{
   const int nk = F.ndim == 3 ? 2 : 1;
   const long long o = mg_index(C, I, J, K);
   for (int m = 0; m < F.nc; m++) {
      const double* r = F.r + m * F.cs;
      double acc = 0.0;
      for (int c = 0; c < nk; c++)
         for (int b = 0; b < 2; b++)
            for (int a = 0; a < 2; a++) acc += r[mg_index(F, 2 * I + a, 2 * J + b, (F.ndim == 3 ? 2 * K : 0) + c)];
      C.f[o + m * C.cs] = acc * (F.ndim == 3 ? 0.125 : 0.25);
      C.u[o + m * C.cs] = 0.0;
   }
}

// the two above in one pass: the coarse cell computes the residuals of its children itself (same
// expressions, same order of accumulation: bit-identical), so r is neither written nor re-read
MG_HD void mg_restrict_residual_cell(const Level& F, const Level& C, int I, int J, int K)
{
   const int nk = F.ndim == 3 ? 2 : 1;
   const long long o = mg_index(C, I, J, K);
   for (int m = 0; m < F.nc; m++) {
      const long long q = m * F.cs;
      double acc = 0.0;
      for (int c = 0; c < nk; c++)
         for (int b = 0; b < 2; b++)
            for (int a = 0; a < 2; a++) {
               const int i = 2 * I + a, j = 2 * J + b, k = (F.ndim == 3 ? 2 * K : 0) + c;
               acc += F.f[mg_index(F, i, j, k) + q] - mg_apply_cell(F, F.u + q, i, j, k);
            }
      C.f[o + m * C.cs] = acc * (F.ndim == 3 ? 0.125 : 0.25);
      C.u[o + m * C.cs] = 0.0;
   }
}

// coarse coefficients of cell (I,J,K): cell fields are the children's mean; the lower face in
// direction a is the mean of the fine faces it covers, divided by 4 (h doubles)
MG_HD void mg_coarsen_cell(const Level& F, const Level& C, int I, int J, int K)
{
   // (constants coarsen on the host: c, m unchanged, d / 4; a stored array is stored on every level)
   const int nk = F.ndim == 3 ? 2 : 1;
   const int k0 = F.ndim == 3 ? 2 * K : 0;
   const double wcell = F.ndim == 3 ? 0.125 : 0.25;
   const long long o = mg_index(C, I, J, K);
   double ac = 0.0, am = 0.0, as = 0.0;
   for (int c = 0; c < nk; c++)
      for (int b = 0; b < 2; b++)
         for (int a = 0; a < 2; a++) {
            const long long of = mg_index(F, 2 * I + a, 2 * J + b, k0 + c);
            if (F.c) ac += F.c[of];
            if (F.m) am += F.m[of];
            if (F.s) as += F.s[of];
         }
   if (C.c) C.c[o] = ac * wcell;
   if (C.m) C.m[o] = am * wcell;
   if (C.s) C.s[o] = as * wcell;
   const double wface = (F.ndim == 3 ? 0.25 : 0.5) * 0.25;
   if (C.d[0]) {
      double a0 = 0.0;
      for (int c = 0; c < nk; c++)
         for (int b = 0; b < 2; b++) a0 += F.d[0][mg_index(F, 2 * I, 2 * J + b, k0 + c)];
      C.d[0][o] = a0 * wface;
   }
   if (C.d[1]) {
      double a1 = 0.0;
      for (int c = 0; c < nk; c++)
         for (int a = 0; a < 2; a++) a1 += F.d[1][mg_index(F, 2 * I + a, 2 * J, k0 + c)];
      C.d[1][o] = a1 * wface;
   }
   if (F.ndim == 3 && C.d[2]) {
      double a2 = 0.0;
      for (int b = 0; b < 2; b++)
         for (int a = 0; a < 2; a++) a2 += F.d[2][mg_index(F, 2 * I + a, 2 * J + b, k0)];
      C.d[2][o] = a2 * wface;
   }
}

// the parent's neighbour on the child's side: periodic image, or the parent itself at a zero-slope boundary
MG_HD int mg_prolong_nb(int I, int upper, int n, int clamp)
{
   if (clamp && ((upper && I + 1 == n) || (!upper && I == 0))) return I;
   return upper ? mg_up(I, n) : mg_dn(I, n);
}

// level 0, after the coefficients are set: a constant D becomes an array entry (fill != 0), and the wrap face of a
// zero-slope direction carries no flux
MG_HD void mg_boundary_faces_cell(const Level& L, int fill, int i, int j, int k)
{
   const long long o = mg_index(L, i, j, k);
   const int idx[3] = {i, j, k};
   for (int a = 0; a < L.ndim; a++) {
      if (!L.d[a]) continue;
      if (fill) L.d[a][o] = L.d_const[a];
      if (L.clamp[a] && idx[a] == 0) L.d[a][o] = 0.0;
   }
}

// fine cell (i,j,k): u_f += cell-centred (bi/tri)linear interpolation of the coarse correction
// (weights 3/4 towards the parent, 1/4 towards the parent's neighbour on the child's side)
MG_HD void mg_prolong_cell(const Level& C, const Level& F, int i, int j, int k)
{
   const int I = i >> 1, J = j >> 1, K = F.ndim == 3 ? (k >> 1) : 0;
   const int I2 = mg_prolong_nb(I, i & 1, C.n[0], C.clamp[0]);
   const int J2 = mg_prolong_nb(J, j & 1, C.n[1], C.clamp[1]);
   for (int m = 0; m < F.nc; m++) {
      const double* cu = C.u + m * C.cs;
      double e;
      if (F.ndim == 3) {
         const int K2 = mg_prolong_nb(K, k & 1, C.n[2], C.clamp[2]);
         const double ea = 0.75 * (0.75 * cu[mg_index(C, I, J, K)] + 0.25 * cu[mg_index(C, I2, J, K)]) +
                           0.25 * (0.75 * cu[mg_index(C, I, J2, K)] + 0.25 * cu[mg_index(C, I2, J2, K)]);
         const double eb = 0.75 * (0.75 * cu[mg_index(C, I, J, K2)] + 0.25 * cu[mg_index(C, I2, J, K2)]) +
                           0.25 * (0.75 * cu[mg_index(C, I, J2, K2)] + 0.25 * cu[mg_index(C, I2, J2, K2)]);
         e = 0.75 * ea + 0.25 * eb;
      } else {
         e = 0.75 * (0.75 * cu[mg_index(C, I, J, 0)] + 0.25 * cu[mg_index(C, I2, J, 0)]) +
             0.25 * (0.75 * cu[mg_index(C, I, J2, 0)] + 0.25 * cu[mg_index(C, I2, J2, 0)]);
      }
      F.u[mg_index(F, i, j, k) + m * F.cs] += e;
   }
}

// ---- finest-level coefficients from SAMRAI-layout PatchData ------------------------------------
// element (i,j,k) of a CellData (axis = -1) or of the `axis` array of a SideData with ghost width
// ng over the box [0, n-1] (pdat_m4arrdim*.i): Fortran order
MG_HD long long mg_samrai_index(const Level& L, int axis, int ng, int i, int j, int k)
{
   const int g2 = L.ndim == 3 ? ng : 0;
   const long long n0 = L.n[0] + 2 * ng + (axis == 0), n1 = L.n[1] + 2 * ng + (axis == 1);
   return (long long)(i + ng) + n0 * ((long long)(j + ng) + n1 * (long long)(k + g2));
}

// scalar block (EllipticFACOps::setM / setC / setD*): a NULL input array = the constant, which the
// caller has put into the level descriptor (m_const, c_const, d_const) with the array pointer null
MG_HD void mg_set_elliptic_cell(const Level& L, const double* m, int ngm, const double* c, int ngc,
                                const double* const* d, const double* const* d2, int ngd, double d_scale,
                                const double* inv_h2, int i, int j, int k)
{
   const long long o = mg_index(L, i, j, k);
   if (L.m) L.m[o] = m[mg_samrai_index(L, -1, ngm, i, j, k)];
   if (L.c) L.c[o] = c[mg_samrai_index(L, -1, ngc, i, j, k)];
   for (int a = 0; a < L.ndim; a++) {
      if (!L.d[a] || !d) continue;  // a constant D stored as an array (zero-slope boundary): mg_boundary_faces_cell
      const long long os = mg_samrai_index(L, a, ngd, i, j, k);
      double D = d[a][os];
      if (d2) D += d2[a][os];
      D *= d_scale;
      L.d[a][o] = D * inv_h2[a];
   }
}

// quaternion block (QuatFACOps::setOperatorCoefficients, QuatFACOps.cc:735-818: sqrt of the mobility;
// QuatLevelSolver::setMatrixCoefficients -> set_j_ij / set_stencil)
MG_HD void mg_set_quat_cell(const Level& L, double gamma, const double* mobility, int ngm,
                            const double* const* face_coef, int ngfc, const double* inv_h2, int i, int j, int k)
{
   const long long o = mg_index(L, i, j, k);
#ifdef __CUDA_ARCH__
   const double sq = ::sqrt(mobility[mg_samrai_index(L, -1, ngm, i, j, k)]);
#else
   const double sq = __builtin_sqrt(mobility[mg_samrai_index(L, -1, ngm, i, j, k)]);
#endif
   L.s[o] = sq;
   L.m[o] = gamma * sq;  // c = 1 is a constant of this block (c_const)
   for (int a = 0; a < L.ndim; a++) L.d[a][o] = face_coef[a][mg_samrai_index(L, a, ngfc, i, j, k)] * inv_h2[a];
}

}  // namespace ampe_mg_cell
