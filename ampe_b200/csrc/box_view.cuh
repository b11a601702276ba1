// SAMRAI-layout box helpers shared by the piecewise kernels (kernels_piecewise.cu, symmetry.cu):
// views of CellData / SideData with ghost widths and a grid-stride launcher over an inclusive
// index box.  Include inside an anonymous namespace after `using namespace ampe;`.
#pragma once
struct Box {
   int ndim;
   int lo[3], hi[3];
};
static Box mkbox(int ndim, const int* lo, const int* hi)
{
   Box b;
   b.ndim = ndim;
   for (int d = 0; d < 3; d++) {
      b.lo[d] = d < ndim ? lo[d] : 0;
      b.hi[d] = d < ndim ? hi[d] : 0;
   }
   return b;
}
// SAMRAI CellData (axis = -1) or one axis of SideData: Fortran order, depth slowest
template <typename T>
struct V {
   T* p;
   int lo0, lo1, lo2, n0, n1;
   long long comp;
   __host__ __device__ T& operator()(int i, int j, int k = 0, int m = 0) const
   {
      return p[(long long)(i - lo0) + (long long)n0 * ((j - lo1) + (long long)n1 * (k - lo2)) +
               comp * m];
   }
   __host__ __device__ V at(int m0) const
   {
      V v = *this;
      v.p = p + comp * m0;
      return v;
   }
};
template <typename T>
static V<T> view(T* p, const Box& b, int axis, int ng)
{
   V<T> v;
   v.p = p;
   int n[3], lo[3];
   for (int d = 0; d < 3; d++) {
      const int g = (d < b.ndim) ? ng : 0;
      lo[d] = b.lo[d] - g;
      n[d] = b.hi[d] - b.lo[d] + 1 + 2 * g + (d == axis ? 1 : 0);
   }
   v.lo0 = lo[0], v.lo1 = lo[1], v.lo2 = lo[2];
   v.n0 = n[0], v.n1 = n[1];
   v.comp = (long long)n[0] * n[1] * n[2];
   return v;
}
typedef V<double> DV;
typedef V<const double> CV;
typedef V<const int> IV;
struct DV3 {
   DV a[3];
};
struct IV3 {
   IV a[3];
};
static DV3 sides(double* const* p, const Box& b, int ng)
{
   DV3 s;
   for (int d = 0; d < b.ndim; d++) s.a[d] = view(p[d], b, d, ng);
   for (int d = b.ndim; d < 3; d++) s.a[d] = s.a[0];
   return s;
}
static DV3 cells3(double* const* p, const Box& b, int ng)
{
   DV3 s;
   for (int d = 0; d < b.ndim; d++) s.a[d] = view(p[d], b, -1, ng);
   for (int d = b.ndim; d < 3; d++) s.a[d] = s.a[0];
   return s;
}
static IV3 isides(const int* const* p, const Box& b, int ng)
{
   IV3 s;
   for (int d = 0; d < b.ndim; d++) s.a[d] = view(p[d], b, d, ng);
   for (int d = b.ndim; d < 3; d++) s.a[d] = s.a[0];
   return s;
}

template <class F>
__global__ void box_kernel(int L0, int L1, int L2, int e0, int e1, long long total, F f)
{
   for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
        t += (long long)gridDim.x * blockDim.x) {
      const int i = (int)(t % e0);
      const int j = (int)((t / e0) % e1);
      const int k = (int)(t / ((long long)e0 * e1));
      f(L0 + i, L1 + j, L2 + k);
   }
}
// run f(i,j,k) over the inclusive box [L,H]
template <class F>
static int for_box(const int* L, const int* H, cudaStream_t st, F f)
{
   const int e0 = H[0] - L[0] + 1, e1 = H[1] - L[1] + 1, e2 = H[2] - L[2] + 1;
   if (e0 <= 0 || e1 <= 0 || e2 <= 0) return AMPE_OK;
   const long long total = (long long)e0 * e1 * e2;
   const int blocks = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
   box_kernel<<<blocks, 256, 0, st>>>(L[0], L[1], L[2], e0, e1, total, f);
   cudaError_t e = cudaGetLastError();
   if (e != cudaSuccess) return ampe_set_err(AMPE_ECUDA, cudaGetErrorString(e));
   return AMPE_OK;
}
// box helpers: cells, sides of axis a (optionally grown by g in the transverse directions)
static void cell_bounds(const Box& b, int g, int* L, int* H)
{
   for (int d = 0; d < 3; d++) {
      const int gg = d < b.ndim ? g : 0;
      L[d] = b.lo[d] - gg;
      H[d] = b.hi[d] + gg;
   }
}
static void side_bounds(const Box& b, int a, int gt, int* L, int* H)
{
   for (int d = 0; d < 3; d++) {
      const int gg = (d < b.ndim && d != a) ? gt : 0;
      L[d] = b.lo[d] - gg;
      H[d] = b.hi[d] + gg + (d == a ? 1 : 0);
   }
}
#define E(a, d) ((a) == (d) ? 1 : 0)
#define ST(stream) ((cudaStream_t)(stream))
