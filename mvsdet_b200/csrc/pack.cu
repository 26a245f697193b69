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

// process_rgb_raw (mvsdet.py:319-333): F.interpolate(rgb[src_id], scale_factor=1/4, 'bilinear'),
// crop to [h,w], laid out as [n, h*w, 3].  With scale 1/4 and align_corners=False the source
// coordinate of output pixel x is 4x + 1.5: the sample is the mean of the 2x2 block at (4x+1, 4y+1),
// evaluated in ATen's order  hl0*(wl0*a + wl1*b) + hl1*(wl0*c + wl1*d)  with all four lambdas 0.5.
__global__ void __launch_bounds__(128) rgb_downsample4_kernel(const float* __restrict__ src,
                                                              const int64_t* __restrict__ ids,
                                                              float* __restrict__ dst, int H, int W,
                                                              int h, int w) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (pix >= h * w) return;
  const int y = pix / w, x = pix - y * w;
  const int64_t v = ids ? ids[n] : n;
  const int y0 = 4 * y + 1, x0 = 4 * x + 1;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* s = src + ((size_t)v * 3 + c) * H * W;
    const float a = __ldg(s + (size_t)y0 * W + x0), b = __ldg(s + (size_t)y0 * W + x1);
    const float cc = __ldg(s + (size_t)y1 * W + x0), d = __ldg(s + (size_t)y1 * W + x1);
    const float top = __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, b));
    const float bot = __fadd_rn(__fmul_rn(0.5f, cc), __fmul_rn(0.5f, d));
    dst[((size_t)n * h * w + pix) * 3 + c] = __fadd_rn(__fmul_rn(0.5f, top), __fmul_rn(0.5f, bot));
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

extern "C" int mvsd_rgb_downsample4(const float* rgb, const int64_t* src_ids, int n_ids, float* out,
                                    int V, int H, int W, int h, int w, void* stream) {
  if (V <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0 || n_ids <= 0)
    return fail(MVSD_ERR_INVALID_ARG, "rgb_downsample4: non-positive dimension");
  if (!rgb || !out) return fail(MVSD_ERR_INVALID_ARG, "rgb_downsample4: null pointer");
  if (h > H / 4 || w > W / 4)
    return fail(MVSD_ERR_INVALID_ARG, "rgb_downsample4: crop [%d,%d] exceeds the quarter-size map of [%d,%d]", h, w, H, W);
  if (n_ids > 65535) return fail(MVSD_ERR_UNSUPPORTED, "rgb_downsample4: %d views > 65535", n_ids);
  dim3 grid((h * w + 127) / 128, n_ids);
  rgb_downsample4_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(rgb, src_ids, out, H, W, h, w);
  count_launch();
  return check_launch("rgb_downsample4");
}
