// Layout helpers: the reference's FPN emits [V,C,H,W] fp32 contiguous
// (projects/NeRF-Det/nerfdet/mvsdet.py:373-376); the kernels of this library
// read channels-last maps so that a bilinear tap / a back-projected pixel is
// one contiguous C-vector.  These two kernels are the drop-in glue for callers
// that cannot run their backbone in torch.channels_last: a tiled transpose
// through shared memory (32x32 tile, +1 padding, coalesced on both sides).
#include "common.cuh"

namespace mvsd {

template <typename TOut>
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, TOut* __restrict__ dst,
                                                   int C, int HW) {
  __shared__ float tile[32][33];
  const int v = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const float* s = src + (size_t)v * C * HW;
  TOut* d = dst + (size_t)v * C * HW;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, pix = p0 + tx;
    tile[r][tx] = (c < C && pix < HW) ? __ldg(s + (size_t)c * HW + pix) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int pix = p0 + r, c = c0 + tx;
    if (c < C && pix < HW) {
      if constexpr (sizeof(TOut) == 2) d[(size_t)pix * C + c] = __float2bfloat16_rn(tile[tx][r]);
      else d[(size_t)pix * C + c] = tile[tx][r];
    }
  }
}

__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     int accumulate, int C, int HW) {
  __shared__ float tile[32][33];
  const int v = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (size_t)v * C * HW;
  float* d = dst + (size_t)v * C * HW;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int pix = p0 + r, c = c0 + tx;
    tile[r][tx] = (c < C && pix < HW) ? __ldg(s + (size_t)pix * C + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, pix = p0 + tx;
    if (c < C && pix < HW) {
      const size_t o = (size_t)c * HW + pix;
      d[o] = accumulate ? d[o] + tile[tx][r] : tile[tx][r];
    }
  }
}

}  // namespace mvsd

using namespace mvsd;

extern "C" int mvsd_pack_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int V, int C,
                                      int H, int W, void* stream) {
  if (V <= 0 || C <= 0 || H <= 0 || W <= 0)
    return fail(MVSD_ERR_INVALID_ARG, "pack: non-positive dimension");
  if (!src || !dst) return fail(MVSD_ERR_INVALID_ARG, "pack: null pointer");
  if (V > 65535) return fail(MVSD_ERR_UNSUPPORTED, "pack: V=%d > 65535", V);
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_dtype == MVSD_F32) pack_kernel<float><<<grid, 256, 0, st>>>(src, static_cast<float*>(dst), C, HW);
  else if (dst_dtype == MVSD_BF16)
    pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), C, HW);
  else return fail(MVSD_ERR_INVALID_ARG, "pack: bad dtype");
  count_launch();
  return check_launch("pack_nchw_to_nhwc");
}

extern "C" int mvsd_unpack_nhwc_to_nchw(const float* src, float* dst, int accumulate, int V, int C,
                                        int H, int W, void* stream) {
  if (V <= 0 || C <= 0 || H <= 0 || W <= 0)
    return fail(MVSD_ERR_INVALID_ARG, "unpack: non-positive dimension");
  if (!src || !dst) return fail(MVSD_ERR_INVALID_ARG, "unpack: null pointer");
  if (V > 65535) return fail(MVSD_ERR_UNSUPPORTED, "unpack: V=%d > 65535", V);
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, V);
  unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, accumulate, C, HW);
  count_launch();
  return check_launch("unpack_nhwc_to_nchw");
}
