// Shared device helpers for the mvsdet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvsdet_b200.h"

#ifndef MVSD_MAX_K
#define MVSD_MAX_K 4      // neighbours per reference view (reference uses 2)
#endif
#define MVSD_MAX_T 8      // top-k hypotheses (reference uses 3)
#define MVSD_MAX_D 64     // depth planes (reference uses 12; sweeps to 64)
#define MVSD_MAX_C 512

namespace mvsd {

// ---- host-side error plumbing (capi.cu) -----------------------------------
int fail(int status, const char* fmt, ...);
int check_launch(const char* what);
void count_launch();
int tuning(int key);

// ---- 4-channel vector access ------------------------------------------------
// Features are read through the read-only path; the big streaming outputs use
// evict-first stores so they do not push the (re-read) feature maps out of L2.
template <typename T> struct Io;

template <> struct Io<float> {
  static __device__ __forceinline__ float4 ld(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
  }
  static __device__ __forceinline__ float4 ld_stream(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
  }
  static __device__ __forceinline__ void st(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
  }
  static __device__ __forceinline__ void st_stream(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
  }
};

__device__ __forceinline__ float4 bf16x4_to_f32(uint2 r) {
  float4 v;
  v.x = __uint_as_float(r.x << 16);
  v.y = __uint_as_float(r.x & 0xffff0000u);
  v.z = __uint_as_float(r.y << 16);
  v.w = __uint_as_float(r.y & 0xffff0000u);
  return v;
}
__device__ __forceinline__ uint2 f32_to_bf16x4(float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}

template <> struct Io<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    return bf16x4_to_f32(__ldg(reinterpret_cast<const uint2*>(p)));
  }
  static __device__ __forceinline__ float4 ld_stream(const __nv_bfloat16* p) {
    return bf16x4_to_f32(__ldcs(reinterpret_cast<const uint2*>(p)));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = f32_to_bf16x4(v);
  }
  static __device__ __forceinline__ void st_stream(__nv_bfloat16* p, float4 v) {
    __stcs(reinterpret_cast<uint2*>(p), f32_to_bf16x4(v));
  }
};

// fp32 vector reduction into global memory (sm_90+): one 16-byte RED instead of
// four scalar atomics.
__device__ __forceinline__ void red_add_f32x4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Deterministic accumulation (the *_det entry points): fp32 REDs add in whatever order the warps
// arrive, so a gradient is reproducible to ~3e-6 of its rms, not bit for bit.  Integer addition is
// associative: every contribution is converted to signed 64-bit fixed point with 32 fractional bits
// (x * 2^32 is exact in fp32, the conversion is exact for |x| < 2^31) and added with 64-bit integer
// REDs; the sum is converted back once (mvsd_fixed_to_float).  Resolution 2.3e-10, range +-2^31.
constexpr float kFixedScale = 4294967296.0f;
__device__ __forceinline__ void red_add_fixed(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * kFixedScale));
}
__device__ __forceinline__ void red_add_fixed4(long long* p, float4 v) {
  red_add_fixed(p + 0, v.x);
  red_add_fixed(p + 1, v.y);
  red_add_fixed(p + 2, v.z);
  red_add_fixed(p + 3, v.w);
}

// ---- plane-sweep sample geometry -------------------------------------------
// One bilinear sample of the homography warp: element offsets (pixel index
// y*W+x, clamped into the map) of the four taps and their weights (zero for
// taps outside the map, i.e. padding_mode='zeros').  any == 0 when every tap
// is outside.
struct alignas(16) WarpSample {
  float w00, w01, w10, w11;     // (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1)
  unsigned p00, p01, p10, p11;  // tap offsets (y*W+x)*scale, valid even when the
                                // weight is 0; p00 == kNoSample <=> no tap at all
};
constexpr unsigned kNoSample = 0xffffffffu;

// Follows mvs_models/module.py:120-142 op by op:
//   q = (rot @ (x,y,1)) * depth + trans          matmul = FMA chain over k
//   px,py = q.xy / q.z                            no z guard, no epsilon
//   g = p / ((size-1)/2) - 1                      the reference's normalisation
//   i = (g + 1) * (size/2) - 0.5                  grid_sample, align_corners=False
// Each step is rounded separately (the reference runs them as separate ATen
// ops), hence the explicit _rn intrinsics: no FMA contraction here.
__device__ __forceinline__ WarpSample make_warp_sample(const float* __restrict__ m,
                                                       float x, float y, float depth,
                                                       int H, int W, int scale) {
  float rx = fmaf(m[2], 1.0f, fmaf(m[1], y, __fmul_rn(m[0], x)));
  float ry = fmaf(m[5], 1.0f, fmaf(m[4], y, __fmul_rn(m[3], x)));
  float rz = fmaf(m[8], 1.0f, fmaf(m[7], y, __fmul_rn(m[6], x)));
  float qx = __fadd_rn(__fmul_rn(rx, depth), m[9]);
  float qy = __fadd_rn(__fmul_rn(ry, depth), m[10]);
  float qz = __fadd_rn(__fmul_rn(rz, depth), m[11]);
  float px = __fdiv_rn(qx, qz);
  float py = __fdiv_rn(qy, qz);
  float gx = __fsub_rn(__fdiv_rn(px, 0.5f * (float)(W - 1)), 1.0f);
  float gy = __fsub_rn(__fdiv_rn(py, 0.5f * (float)(H - 1)), 1.0f);
  float ix = __fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f * (float)W), 0.5f);
  float iy = __fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f * (float)H), 0.5f);
  float x0f = floorf(ix), y0f = floorf(iy);
  WarpSample s;
  // NaN / inf / far-away coordinates fail these comparisons -> no tap.
  bool inx = (x0f >= -1.0f) && (x0f <= (float)(W - 1));
  bool iny = (y0f >= -1.0f) && (y0f <= (float)(H - 1));
  if (!(inx && iny)) {
    s.w00 = s.w01 = s.w10 = s.w11 = 0.f;
    s.p00 = s.p01 = s.p10 = s.p11 = kNoSample;
    return s;
  }
  int x0 = (int)x0f, y0 = (int)y0f;
  float wx = __fsub_rn(ix, x0f), wy = __fsub_rn(iy, y0f);
  float ex = __fsub_rn(1.0f, wx), ey = __fsub_rn(1.0f, wy);
  bool vx0 = x0 >= 0, vx1 = x0 + 1 <= W - 1;
  bool vy0 = y0 >= 0, vy1 = y0 + 1 <= H - 1;
  int cx0 = vx0 ? x0 : 0, cx1 = vx1 ? x0 + 1 : W - 1;
  int cy0 = vy0 ? y0 : 0, cy1 = vy1 ? y0 + 1 : H - 1;
  s.w00 = (vx0 && vy0) ? __fmul_rn(ey, ex) : 0.f;
  s.w01 = (vx1 && vy0) ? __fmul_rn(ey, wx) : 0.f;
  s.w10 = (vx0 && vy1) ? __fmul_rn(wy, ex) : 0.f;
  s.w11 = (vx1 && vy1) ? __fmul_rn(wy, wx) : 0.f;
  s.p00 = (unsigned)((cy0 * W + cx0) * scale);
  s.p01 = (unsigned)((cy0 * W + cx1) * scale);
  s.p10 = (unsigned)((cy1 * W + cx0) * scale);
  s.p11 = (unsigned)((cy1 * W + cx1) * scale);
  return s;
}

// ---- voxel projection test (back-projection) --------------------------------
// Follows backproject_Weigh, mvsdet.py:1384-1391 and :1395-1427
// (SURVEY.md Appendix A.3).  torch.bmm with K=4 accumulates as an FMA chain in
// k order (checked against ATen on CPU, bit for bit); round() is half-to-even.
struct VoxelHit {
  int pix;        // y*w + x inside the [h,w] crop, -1 when out of bounds
  int x, y;
  int jstar;      // hypothesis that supplies the weight (first maximum), -1 if none
  float weight;   // max_j (pass_j ? pn_j : 0)
  bool valid;     // any_j pass_j
};

__device__ __forceinline__ bool project_voxel(const float* __restrict__ P, float X, float Y,
                                              float Z, int h, int w, int& xi, int& yi,
                                              float& z) {
  float p0 = fmaf(P[3], 1.0f, fmaf(P[2], Z, fmaf(P[1], Y, __fmul_rn(P[0], X))));
  float p1 = fmaf(P[7], 1.0f, fmaf(P[6], Z, fmaf(P[5], Y, __fmul_rn(P[4], X))));
  float p2 = fmaf(P[11], 1.0f, fmaf(P[10], Z, fmaf(P[9], Y, __fmul_rn(P[8], X))));
  float xr = rintf(__fdiv_rn(p0, p2));
  float yr = rintf(__fdiv_rn(p1, p2));
  z = p2;
  bool inb = (xr >= 0.f) && (yr >= 0.f) && (xr < (float)w) && (yr < (float)h) && (p2 > 0.f);
  xi = inb ? (int)xr : 0;
  yi = inb ? (int)yr : 0;
  return inb;
}

template <int TMAX>
__device__ __forceinline__ VoxelHit test_voxel(const float* __restrict__ P, float X, float Y,
                                               float Z, const float* __restrict__ depth,
                                               const float* __restrict__ prob, int64_t sy,
                                               int64_t sx, int64_t st, float vs_z, int h, int w,
                                               int T) {
  VoxelHit hit;
  float z;
  bool inb = project_voxel(P, X, Y, Z, h, w, hit.x, hit.y, z);
  hit.pix = -1;
  hit.jstar = -1;
  hit.weight = 0.f;
  hit.valid = false;
  if (!inb) return hit;
  hit.pix = hit.y * w + hit.x;
  const int64_t base = (int64_t)hit.y * sy + (int64_t)hit.x * sx;
  float pr[TMAX];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < TMAX; ++j) {
    if (j < T) {
      pr[j] = __ldg(prob + base + j * st);
      sum = (j == 0) ? pr[j] : __fadd_rn(sum, pr[j]);
    }
  }
  // weight = max over the T candidates (pass ? pn : 0); torch.max returns the
  // first maximal index, which is where its backward routes the gradient.
  float best = 0.f;
  int bestj = -1;
  bool bestpass = false;
#pragma unroll
  for (int j = 0; j < TMAX; ++j) {
    if (j < T) {
      float d = __ldg(depth + base + j * st);
      bool pass = (z > __fsub_rn(d, vs_z)) && (z < __fadd_rn(d, vs_z));
      float cand = pass ? __fdiv_rn(pr[j], sum) : 0.f;
      hit.valid |= pass;
      if (j == 0 || cand > best) {
        best = cand;
        bestj = j;
        bestpass = pass;
      }
    }
  }
  hit.weight = best;
  hit.jstar = bestpass ? bestj : -1;
  return hit;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mvsd
