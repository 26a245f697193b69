// Layout helpers: the reference's FPN emits [V,C,H,W] fp32 contiguous
// (projects/NeRF-Det/nerfdet/mvsdet.py:373-376); the kernels of this library
// read channels-last maps so that a bilinear tap / a back-projected pixel is
// one contiguous C-vector.  These two kernels are the drop-in glue for callers
// that cannot run their backbone in torch.channels_last: a tiled transpose
// through shared memory (64x64 tile, +1 padding, 128 B per warp access on both sides).
#include "common.cuh"

namespace mvsd {

constexpr int kPT = 64;      // tile edge: 64 pixels x 64 channels per CTA (16 KB in flight per CTA)

// 64x64 tiles, 256 threads, 16 loads + 16 stores per thread: 4x the bytes in
// flight of the first 32x32 version (3.6 -> measured in profiles/).  Row stride
// 65 keeps the transposed shared-memory reads conflict-free (2-way for the
// bf16 pair reads).
template <typename TOut>
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, TOut* __restrict__ dst,
                                                   int C, int HW) {
  __shared__ float tile[kPT][kPT + 1];                    // [channel][pixel]
  const int v = blockIdx.z;
  const int p0 = blockIdx.x * kPT, c0 = blockIdx.y * kPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* s = src + (size_t)v * C * HW;
  TOut* d = dst + (size_t)v * C * HW;
  // read: a warp reads 32 consecutive pixels of one channel (128 B)
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int idx = it * 8 + warp;                        // 0..127: (channel row, pixel half)
    const int r = idx >> 1, px = (idx & 1) * 32 + lane;
    const int c = c0 + r, pix = p0 + px;
    tile[r][px] = (c < C && pix < HW) ? __ldg(s + (size_t)c * HW + pix) : 0.f;
  }
  __syncthreads();
  // write: a warp writes the 64 channels of one pixel (lane = channel pair)
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int px = it * 8 + warp;
    const int pix = p0 + px, c = c0 + 2 * lane;
    if (pix >= HW || c >= C) continue;
    const float a = tile[2 * lane][px], b = tile[2 * lane + 1][px];
    const bool pair = (c + 1 < C) && !(C & 1);            // odd C: pixel rows are not pair-aligned
    if constexpr (sizeof(TOut) == 2) {
      if (pair) {
        *reinterpret_cast<__nv_bfloat162*>(d + (size_t)pix * C + c) = __floats2bfloat162_rn(a, b);
      } else {
        d[(size_t)pix * C + c] = __float2bfloat16_rn(a);
        if (c + 1 < C) d[(size_t)pix * C + c + 1] = __float2bfloat16_rn(b);
      }
    } else {
      if (pair) {
        *reinterpret_cast<float2*>(d + (size_t)pix * C + c) = make_float2(a, b);
      } else {
        d[(size_t)pix * C + c] = a;
        if (c + 1 < C) d[(size_t)pix * C + c + 1] = b;
      }
    }
  }
}

__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     int accumulate, int C, int HW) {
  __shared__ float tile[kPT][kPT + 1];                    // [pixel][channel]
  const int v = blockIdx.z;
  const int p0 = blockIdx.x * kPT, c0 = blockIdx.y * kPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* s = src + (size_t)v * C * HW;
  float* d = dst + (size_t)v * C * HW;
  // read: a warp reads 32 consecutive channels of one pixel (128 B)
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int idx = it * 8 + warp;
    const int px = idx >> 1, ch = (idx & 1) * 32 + lane;
    const int pix = p0 + px, c = c0 + ch;
    tile[px][ch] = (c < C && pix < HW) ? __ldcs(s + (size_t)pix * C + c) : 0.f;
  }
  __syncthreads();
  // write: a warp writes 32 consecutive pixels of one channel (128 B)
#pragma unroll
  for (int it = 0; it < 16; ++it) {
    const int idx = it * 8 + warp;
    const int ch = idx >> 1, px = (idx & 1) * 32 + lane;
    const int c = c0 + ch, pix = p0 + px;
    if (c < C && pix < HW) {
      const size_t o = (size_t)c * HW + pix;
      d[o] = accumulate ? d[o] + tile[px][ch] : tile[px][ch];
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
  dim3 grid((HW + kPT - 1) / kPT, (C + kPT - 1) / kPT, V);
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
  dim3 grid((HW + kPT - 1) / kPT, (C + kPT - 1) / kPT, V);
  unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, accumulate, C, HW);
  count_launch();
  return check_launch("unpack_nhwc_to_nchw");
}
