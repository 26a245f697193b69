// Layout helpers: the reference's FPN emits [V,C,H,W] fp32 contiguous
// (projects/NeRF-Det/nerfdet/mvsdet.py:373-376); the kernels of this library
// read channels-last maps so that a bilinear tap / a back-projected pixel is
// one contiguous C-vector.  These two kernels are the drop-in glue for callers
// that cannot run their backbone in torch.channels_last: a tiled transpose
// through shared memory (64x64 tile, +1 padding, 128 B per warp access on both sides).
#include "common.cuh"

namespace mvsd {

__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const unsigned*>(&v);
}

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

// ---- 16-byte variants (the shipped FPN shapes: HW % 4 == 0, C % 8 == 0, 16-byte aligned bases) ------
// Same 64 x 64 tile, but every global access moves 16 bytes per lane on BOTH sides (the kernels above
// move 4 bytes per lane on the transposed side: 16 loads + 8..16 stores per thread): 4 float4 loads
// and 2 x 16-byte stores per thread for the pack, 4 + 4 for the unpack.  Shared memory stays
// scalar with the odd row stride (2-way bank conflicts at most, far from the 128 B/clk limit).
template <typename TOut>
__global__ void __launch_bounds__(256) pack16_kernel(const float* __restrict__ src, TOut* __restrict__ dst,
                                                     int C, int HW) {
  __shared__ float tile[kPT][kPT + 1];                    // [channel][pixel]
  const int v = blockIdx.z;
  const int p0 = blockIdx.x * kPT, c0 = blockIdx.y * kPT;
  const int t = threadIdx.x;
  const float* s = src + (size_t)v * C * HW;
  TOut* d = dst + (size_t)v * C * HW;
#pragma unroll
  for (int it = 0; it < 4; ++it) {                        // 64 channels x 16 pixel quads
    const int idx = it * 256 + t;
    const int r = idx >> 4, q = idx & 15;
    const int c = c0 + r, pix = p0 + 4 * q;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C && pix < HW) val = __ldcs(reinterpret_cast<const float4*>(s + (size_t)c * HW + pix));   // HW % 4 == 0
    tile[r][4 * q + 0] = val.x; tile[r][4 * q + 1] = val.y;
    tile[r][4 * q + 2] = val.z; tile[r][4 * q + 3] = val.w;
  }
  __syncthreads();
  if constexpr (sizeof(TOut) == 2) {                      // 64 pixels x 8 groups of 8 channels (16 B of bf16)
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = it * 256 + t;
      const int g = idx >> 2 & 7, px = (idx & 3) + 4 * (idx >> 5);
      const int pix = p0 + px, c = c0 + 8 * g;
      if (pix >= HW || c >= C) continue;                  // C % 8 == 0: a group is all in or all out
      uint4 o;
      o.x = pack_bf16x2(tile[8 * g + 0][px], tile[8 * g + 1][px]);
      o.y = pack_bf16x2(tile[8 * g + 2][px], tile[8 * g + 3][px]);
      o.z = pack_bf16x2(tile[8 * g + 4][px], tile[8 * g + 5][px]);
      o.w = pack_bf16x2(tile[8 * g + 6][px], tile[8 * g + 7][px]);
      *reinterpret_cast<uint4*>(d + (size_t)pix * C + c) = o;
    }
  } else {                                                // 64 pixels x 16 groups of 4 channels
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 256 + t;
      const int g = idx >> 2 & 15, px = (idx & 3) + 4 * (idx >> 6);
      const int pix = p0 + px, c = c0 + 4 * g;
      if (pix >= HW || c >= C) continue;
      *reinterpret_cast<float4*>(d + (size_t)pix * C + c) =
          make_float4(tile[4 * g + 0][px], tile[4 * g + 1][px], tile[4 * g + 2][px], tile[4 * g + 3][px]);
    }
  }
}

__global__ void __launch_bounds__(256) unpack16_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                       int accumulate, int C, int HW) {
  __shared__ float tile[kPT][kPT + 1];                    // [pixel][channel]
  const int v = blockIdx.z;
  const int p0 = blockIdx.x * kPT, c0 = blockIdx.y * kPT;
  const int t = threadIdx.x;
  const float* s = src + (size_t)v * C * HW;
  float* d = dst + (size_t)v * C * HW;
#pragma unroll
  for (int it = 0; it < 4; ++it) {                        // 64 pixels x 16 channel quads
    const int idx = it * 256 + t;
    const int px = idx >> 4, q = idx & 15;
    const int pix = p0 + px, c = c0 + 4 * q;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix < HW && c < C) val = __ldcs(reinterpret_cast<const float4*>(s + (size_t)pix * C + c));    // C % 4 == 0
    tile[px][4 * q + 0] = val.x; tile[px][4 * q + 1] = val.y;
    tile[px][4 * q + 2] = val.z; tile[px][4 * q + 3] = val.w;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; ++it) {                        // 64 channels x 16 pixel quads
    const int idx = it * 256 + t;
    const int ch = idx >> 4, q = idx & 15;
    const int c = c0 + ch, pix = p0 + 4 * q;
    if (c >= C || pix >= HW) continue;
    float4 val = make_float4(tile[4 * q + 0][ch], tile[4 * q + 1][ch], tile[4 * q + 2][ch], tile[4 * q + 3][ch]);
    float4* o = reinterpret_cast<float4*>(d + (size_t)c * HW + pix);
    if (accumulate) {
      const float4 old = *o;
      val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
    }
    *o = val;
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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
  const bool wide = HW % 4 == 0 && C % 8 == 0 && aligned16(src) && aligned16(dst);
  if (dst_dtype == MVSD_F32) {
    if (wide) pack16_kernel<float><<<grid, 256, 0, st>>>(src, static_cast<float*>(dst), C, HW);
    else pack_kernel<float><<<grid, 256, 0, st>>>(src, static_cast<float*>(dst), C, HW);
  } else if (dst_dtype == MVSD_BF16) {
    if (wide) pack16_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), C, HW);
    else pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), C, HW);
  } else {
    return fail(MVSD_ERR_INVALID_ARG, "pack: bad dtype");
  }
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
  if (HW % 4 == 0 && C % 4 == 0 && aligned16(src) && aligned16(dst))
    unpack16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, accumulate, C, HW);
  else
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
