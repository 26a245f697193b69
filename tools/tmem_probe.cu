// Probe: TMEM as a warp-private fp32 accumulator store from ordinary CUDA warps
// (tcgen05.alloc / st / ld / dealloc with the 32x32b shape), as used by the
// block-merging plane-sweep backward for its per-pixel reference gradients.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(taddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float& a, float& b, float& c, float& d) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int COLS>
__global__ void __launch_bounds__(128) probe(float* out, int iters) {
  __shared__ uint32_t s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&s_base)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = s_base + ((uint32_t)(warp * 32) << 16);   // this warp's lane quarter
  // zero all cells
  for (int c = 0; c < COLS; c += 4) tmem_st4(base + c, 0.f, 0.f, 0.f, 0.f);
  tmem_wait_st();
  // accumulate: cell i gets += (i+1)*(lane+1) + warp, iters times, dynamic column index
  for (int it = 0; it < iters; ++it) {
    for (int i = 0; i < COLS / 4; ++i) {
      const int cell = (i * 7 + it) % (COLS / 4);          // data-dependent order
      float a, b, c, d;
      tmem_ld4(base + 4 * cell, a, b, c, d);
      tmem_wait_ld();
      const float v = (float)((cell + 1) * (lane + 1) + warp);
      tmem_st4(base + 4 * cell, a + v, b + 2 * v, c + 3 * v, d + 4 * v);
    }
    tmem_wait_st();
  }
  float acc = 0.f;
  for (int cell = 0; cell < COLS / 4; ++cell) {
    float a, b, c, d;
    tmem_ld4(base + 4 * cell, a, b, c, d);
    tmem_wait_ld();
    const float v = (float)((cell + 1) * (lane + 1) + warp) * iters;
    acc += fabsf(a - v) + fabsf(b - 2 * v) + fabsf(c - 3 * v) + fabsf(d - 4 * v);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_base), "n"(COLS) : "memory");
}

int main() {
  const int grid = 148 * 8, iters = 64;
  float* out;
  cudaMalloc(&out, grid * 128 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    probe<128><<<grid, 128>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s, %.3f ms\n", cudaGetErrorString(e), ms);
  if (e != cudaSuccess) return 1;
  float* h = new float[grid * 128];
  cudaMemcpy(h, out, grid * 128 * sizeof(float), cudaMemcpyDeviceToHost);
  double tot = 0;
  for (int i = 0; i < grid * 128; ++i) tot += h[i];
  // RMW traffic: grid*4 warps * iters * 32 cells * 512 B read + 512 B write
  const double bytes = (double)grid * 4 * iters * 32 * 1024.0;
  printf("total abs error %.1f (want 0); RMW rate %.1f GB/s chip (ld+st bytes), %.2f B/clk/SM @1.9GHz\n", tot,
         bytes / ms / 1e6, bytes / ms / 1e6 / 148 / 1.9);
  return tot != 0.0;
}
