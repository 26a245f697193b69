// Per-SM issue rates of the instructions the sweep kernels are made of (B200):
// FFMA (3-reg), FFMA2 (f32x2), FMUL2, LOP3, SHF, PRMT, IMAD.U32 (shift on the fma
// pipe), and mixes.  Each kernel runs ILP independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_pipes microbench_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, unsigned useed) {
  float a[ILP], b = seed, c = seed * 0.5f;
  u64 p[ILP], pb, pc;
  unsigned u[ILP], ub = useed;
  asm("mov.b64 %0, {%1,%2};" : "=l"(pb) : "f"(b), "f"(b));
  asm("mov.b64 %0, {%1,%2};" : "=l"(pc) : "f"(c), "f"(c));
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = threadIdx.x + i; asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i])); u[i] = threadIdx.x * 2654435761u + i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) a[i] = fmaf(a[i], b, c);                                                       // FFMA
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc)); // FFMA2
      if (MODE == 2) asm volatile("lop3.b32 %0, %0, %1, 0xffff0000, 0x6a;" : "+r"(u[i]) : "r"(ub));   // LOP3
      if (MODE == 3) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(u[i]) : "r"(ub));            // SHF
      if (MODE == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u[i]) : "r"(ub));             // PRMT
      if (MODE == 5) asm volatile("mad.lo.u32 %0, %0, 0x10000, %1;" : "+r"(u[i]) : "r"(ub));          // IMAD
      if (MODE == 6) {   // mix: 2 LOP3/SHF + 1 FFMA2 (the unpack+blend ratio: 4 unpack per 2 FFMA2)
        asm volatile("lop3.b32 %0, %0, %1, 0xffff0000, 0x6a;" : "+r"(u[i]) : "r"(ub));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(u[i]) : "r"(ub));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
      }
      if (MODE == 7) {   // mix: 1 LOP3 + 1 IMAD.U32-shift + 1 FFMA2
        asm volatile("lop3.b32 %0, %0, %1, 0xffff0000, 0x6a;" : "+r"(u[i]) : "r"(ub));
        asm volatile("mad.lo.u32 %0, %0, 0x10000, %1;" : "+r"(u[i]) : "r"(ub));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pb), "l"(pc));
      }
      if (MODE == 8) {   // scalar equivalent: 2 LOP3/SHF + 2 FFMA
        asm volatile("lop3.b32 %0, %0, %1, 0xffff0000, 0x6a;" : "+r"(u[i]) : "r"(ub));
        asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(u[i]) : "r"(ub));
        a[i] = fmaf(a[i], b, c);
        a[i] = fmaf(a[i], c, b);
      }
    }
  }
  float s = 0; unsigned us = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p[i])); s += a[i] + x + y; us += u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + us;
}

template <int MODE>
void run(const char* name, int per_iter, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8;
  float ms = 0;
  for (int r = 0; r < 2; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 1.0001f, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  const double winst = (double)grid * 8 * ITERS * ILP * per_iter;     // warp-instructions
  const double per_sm_clk = winst / (ms * 1e-3) / 148 / 1.965e9;
  printf("%-28s %7.3f ms  %5.2f warp-instr/clk/SM (%4.2f per SMSP)\n", name, ms, per_sm_clk, per_sm_clk / 4);
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("FFMA", 1, out); run<1>("FFMA2", 1, out); run<2>("LOP3", 1, out); run<3>("SHF", 1, out);
  run<4>("PRMT", 1, out); run<5>("IMAD.U32", 1, out);
  run<6>("LOP3+SHF+FFMA2", 3, out); run<7>("LOP3+IMAD+FFMA2", 3, out); run<8>("LOP3+SHF+2FFMA", 4, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
