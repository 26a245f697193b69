// Micro-benchmark behind DESIGN.md "plane-sweep backward: what bounds the scatter".
// Measures, on the GPU it runs on, the payload rate of fp32 reductions into an
// L2-resident 98 MB buffer (the size of one scene's g_feat) issued three ways:
//   red_lsu   red.global.add.v4.f32 from registers (what the kernels use)
//   red_bulk  cp.reduce.async.bulk.global.shared::cta.add.f32 of 1 KB smem rows (TMA)
//   st_lsu    plain st.global.v4.f32 to the same addresses (egress without the L2 ALU)
// Addresses are pseudo-random 1 KB cells (one pixel x 256 channels), like the
// bilinear scatter.   nvcc -arch=sm_100a -O3 -o microbench_red microbench_red.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int MODE>   // 0 red, 1 store
__global__ void __launch_bounds__(128) k_lsu(float* buf, unsigned cells, int iters) {
  const int lane = threadIdx.x & 31;
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int i = 0; i < iters; ++i) {
    const unsigned cell = hash32(gw * 7919u + i) % cells;
    float* p = buf + (size_t)cell * 256 + 4 * lane;
    if (MODE == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p + 128), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else {
      *reinterpret_cast<float4*>(p) = v;
      *reinterpret_cast<float4*>(p + 128) = v;
    }
  }
}

// neighbouring-cell variant: consecutive iterations hit consecutive cells (row runs)
__global__ void __launch_bounds__(128) k_lsu_seq(float* buf, unsigned cells, int iters) {
  const int lane = threadIdx.x & 31;
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  unsigned cell = hash32(gw * 7919u) % cells;
  for (int i = 0; i < iters; ++i) {
    if ((i & 7) == 0) cell = hash32(gw * 7919u + i) % cells;
    cell = (cell + 1) % cells;
    float* p = buf + (size_t)cell * 256 + 4 * lane;
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p + 128), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
}

__global__ void __launch_bounds__(128) k_bulk(float* buf, unsigned cells, int iters, int rows) {
  extern __shared__ __align__(128) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float* mine = sm + (size_t)warp * rows * 256;
  for (int i = lane; i < rows * 256; i += 32) mine[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    for (int i = 0; i < iters; i += rows) {
      unsigned cell = hash32(gw * 7919u + i) % (cells - rows);
      float* g = buf + (size_t)cell * 256;
      unsigned s = (unsigned)__cvta_generic_to_shared(mine);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                   :: "l"(g), "r"(s), "r"(rows * 1024) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const unsigned cells = 96000;                 // 20 views x 4800 pixels
  float* buf;
  cudaMalloc(&buf, (size_t)cells * 1024);
  cudaMemset(buf, 0, (size_t)cells * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2048;
  for (int cps = 1; cps <= 8; cps *= 2) {
    const int grid = sms * cps;
    const double bytes = (double)grid * 4 * iters * 1024.0;
    float ms;
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_lsu<0><<<grid, 128>>>(buf, cells, iters);
        else if (mode == 1) k_lsu<1><<<grid, 128>>>(buf, cells, iters);
        else k_lsu_seq<<<grid, 128>>>(buf, cells, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("%-8s ctas/sm=%d warps/sm=%2d  %8.1f GB/s payload  (%.3f ms)\n",
             mode == 0 ? "red_lsu" : mode == 1 ? "st_lsu" : "red_seq", cps, cps * 4, bytes / ms / 1e6, ms);
    }
    for (int rows = 1; rows <= 8; rows *= 2) {
      if (cps * 4 * rows * 1024 > 200 * 1024) continue;
      cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * rows * 1024);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_bulk<<<grid, 128, 4 * rows * 1024>>>(buf, cells, iters, rows);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("red_bulk ctas/sm=%d warps/sm=%2d rows=%d  %8.1f GB/s payload  (%.3f ms)\n", cps, cps * 4, rows,
             bytes / ms / 1e6, ms);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
