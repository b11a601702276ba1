// FP64 issue-rate probe for B200 (sm_100a): what fraction of the nominal 64 FP64 lanes / SM / clock do dependent
// chains of DFMA / DADD / DMUL reach, as a function of the independent chains per thread (ILP) and of the resident
// warps per SM?  The fused RHS kernels sit at 54-61 % "FP64 pipe" in ncu with 16-32 warps per SM; this says what
// the pipe itself can sustain and how many warps x chains it needs (profiles/README.md, "FP64 roof").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/probe/fp64_peak tools/probe/fp64_peak.cu
#include <cuda_runtime.h>

#include <cstdio>

template <int ILP, int OP>
__global__ void chain(double* out, double a, double b, int iters)
{
   double x[ILP];
#pragma unroll
   for (int i = 0; i < ILP; i++) x[i] = a + i + threadIdx.x * 1e-9;
   for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
         for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = fma(x[i], b, a);
            if (OP == 1) x[i] = x[i] + b;
            if (OP == 2) x[i] = x[i] * b;
            if (OP == 3) x[i] = (u % 3 == 0) ? fma(x[i], b, a) : ((u % 3 == 1) ? x[i] + b : x[i] * b);
         }
      }
   }
   double s = 0.0;
#pragma unroll
   for (int i = 0; i < ILP; i++) s += x[i];
   if (s == 12345.678) out[0] = s;
}

template <int ILP, int OP>
void run(const char* name, int sms, double clk_ghz, double* d)
{
   const int iters = 4000;
   for (int warps : {4, 8, 16, 24, 32, 48, 64}) {
      const int threads = 256, blocks_per_sm = warps * 32 / threads;
      if (blocks_per_sm < 1) {
         continue;
      }
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      chain<ILP, OP><<<sms * blocks_per_sm, threads>>>(d, 1.0000001, 0.9999999, 10);
      cudaEventRecord(e0);
      chain<ILP, OP><<<sms * blocks_per_sm, threads>>>(d, 1.0000001, 0.9999999, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double inst = (double)sms * blocks_per_sm * threads * iters * 8.0 * ILP;  // thread-level FP64 instructions
      const double per_clk_sm = inst / (ms * 1e-3) / (clk_ghz * 1e9) / sms;
      printf("%-6s ILP %d  warps/SM %2d  %.1f FP64 lane-inst/clk/SM  (%.0f %% of 64)\n", name, ILP, warps, per_clk_sm,
             100.0 * per_clk_sm / 64.0);
   }
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   int clk_khz = 0;
   cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
   const double clk = clk_khz * 1e-6;
   printf("%s, %d SMs, clock %.3f GHz (max; the achieved clock is what nvidia-smi shows under load)\n", p.name,
          p.multiProcessorCount, clk);
   double* d;
   cudaMalloc(&d, 8);
   run<1, 0>("DFMA", p.multiProcessorCount, clk, d);
   run<2, 0>("DFMA", p.multiProcessorCount, clk, d);
   run<4, 0>("DFMA", p.multiProcessorCount, clk, d);
   run<8, 0>("DFMA", p.multiProcessorCount, clk, d);
   run<4, 1>("DADD", p.multiProcessorCount, clk, d);
   run<4, 2>("DMUL", p.multiProcessorCount, clk, d);
   run<1, 3>("MIX", p.multiProcessorCount, clk, d);
   run<4, 3>("MIX", p.multiProcessorCount, clk, d);
   cudaFree(d);
   return 0;
}
