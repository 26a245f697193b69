// Shared pieces of the plane-sweep kernels (forward: plane_sweep_fwd.cu,
// backward: plane_sweep_bwd.cu), sm_100a.
//
// Replaces projects/NeRF-Det/nerfdet/mvsdet.py:439-467 and
// mvs_models/module.py:105-146 of the reference: instead of materialising the
// repeated reference volume, k warped [V,C,D,H,W] volumes, their squares and
// running sums (~30 passes over 1.18 GB tensors), one kernel gathers the
// bilinear taps of the k neighbours and writes the variance once.
//
// Work decomposition ("runs").  Features are channels-last ([V,H,W,C]), so a
// bilinear tap is a contiguous C-vector.  A warp owns a horizontal run of
// kRun = 8 pixels of one row of one reference view, for one slice of 128*G
// channels (lane l <-> channels slice*128G + 128g + 4l .. +3 for g < G: 16-byte
// vectors, every request a coalesced 512 bytes), and walks the D planes in its
// outer loop and the 8 pixels in its inner loop:
//
//   * sample geometry (homography, divides, floor, tap weights) is
//     warp-uniform; it is computed ONCE per (pixel, plane, neighbour) with one
//     lane per sample -- 32 samples per pass -- parked in shared memory and
//     re-read as a broadcast, so the inner loop is loads and FMAs only;
//   * going left to right along the run, the right tap column of pixel x is
//     the left tap column of pixel x+1 whenever the source position advances
//     by one pixel (the common case for pose-space neighbours).  The forward
//     keeps those two C-vectors in registers (half the L1 traffic); the
//     backward sums the two contributions to the shared column in registers
//     and emits ONE vector RED instead of two;
//   * the 4 warps of a CTA take 4 consecutive rows, so the vertical overlap of
//     their footprints is served by L1.
//
// HBM-bound by design: no contraction, no tensor cores.
#pragma once
#include "common.cuh"

namespace mvsd {

constexpr int kRun = 8;                    // pixels per warp run
constexpr int kRows = 4;                   // rows (= warps) per CTA
constexpr int kSweepThreads = kRows * 32;
constexpr unsigned kNoTap = 0xfffffffeu;   // "nothing cached / nothing open"

struct SweepParams {
  const void* feat;        // nhwc features
  const int32_t* nbr;      // [V,k] or nullptr (warp-only: source = same index)
  const float* hom;        // [V,k,12]
  const float* depth;      // [V,D]
  void* out;               // fwd output [V,D,H,W,C]
  const void* g_out;       // bwd upstream gradient, same layout
  float* g_feat;           // bwd: nhwc fp32, accumulated with RED
  int V, C, D, H, W, k;
  int ref_begin;           // feat index of reference view 0 (view sharding)
  int runs_x, tiles_y, slices;
  int pf;                  // tap prefetch distance in run steps (0 = off)
};

struct SweepCoord {
  int v, y, x0, npix, c0;
};

template <int G>
__device__ __forceinline__ SweepCoord sweep_coord(const SweepParams& p, int warp, int lane) {
  SweepCoord c;
  int t = blockIdx.x;
  const int xr = t % p.runs_x; t /= p.runs_x;
  const int yt = t % p.tiles_y; t /= p.tiles_y;
  const int slice = t % p.slices;
  c.v = t / p.slices;
  c.y = yt * kRows + warp;
  c.x0 = xr * kRun;
  c.npix = min(kRun, p.W - c.x0);
  c.c0 = slice * 128 * G + 4 * lane;
  return c;
}

// One pass of sample geometry: lane s computes sample
//   (plane d0 + s / (kRun*k), pixel x0 + (s % (kRun*k)) / k, neighbour s % k).
__device__ __forceinline__ void fill_samples(WarpSample* tab, const SweepParams& p,
                                             const SweepCoord& c, int d0, int ppf, int lane) {
  const int k = p.k, spp = kRun * k;
  if (lane < ppf * spp) {
    const int dd = lane / spp, rem = lane - dd * spp;
    const int i = rem / k, j = rem - i * k;
    const int d = d0 + dd;
    WarpSample s;
    s.w00 = s.w01 = s.w10 = s.w11 = 0.f;
    s.p00 = s.p01 = s.p10 = s.p11 = kNoSample;
    if (d < p.D && i < c.npix) {
      const float* m = p.hom + ((size_t)c.v * k + j) * 12;
      float mm[12];
#pragma unroll
      for (int t = 0; t < 12; ++t) mm[t] = __ldg(m + t);
      s = make_warp_sample(mm, (float)(c.x0 + i), (float)c.y, __ldg(p.depth + (size_t)c.v * p.D + d),
                           p.H, p.W, p.C);
    }
    tab[lane] = s;
  }
}

__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4scale(float4 a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 f4fma(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z),
                     fmaf(a.w, b.w, c.w));
}

// 64-bit address = base + 32-bit unsigned element offset: one IMAD.WIDE.U32.
template <typename T>
__device__ __forceinline__ const T* at(const T* base, unsigned off) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) +
                                    (unsigned long long)off * sizeof(T));
}
template <typename T>
__device__ __forceinline__ T* at(T* base, unsigned off) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + (unsigned long long)off * sizeof(T));
}

// lane-activity of channel group g of this warp's slice
template <bool FULL>
__device__ __forceinline__ bool group_on(int c0, int g, int C) {
  return FULL || (c0 + 128 * g < C);
}

// The bilinear sample of one neighbour, nw*w00 + ne*w01 + sw*w10 + se*w11 (the
// order ATen's sampler uses), with the register-resident column cache:
//   col[slot][0|1][g]: slot (step & 1) receives the RIGHT column loaded at this
//   step; slot 1 - (step & 1) holds the previous step's right column, which is
//   this step's LEFT column when id_top / id_bot (offsets of what those
//   registers hold) match.  CACHE = false always loads all four taps.
template <typename TIn, int G, bool FULL, bool CACHE>
__device__ __forceinline__ void gather_taps(const TIn* __restrict__ src, const WarpSample& s,
                                            int c0, int C, float4 (&col)[2][2][G],
                                            unsigned& id_top, unsigned& id_bot, int step,
                                            float4 (&wv)[G]) {
  const int r = CACHE ? (step & 1) : 0, l = r ^ 1;
  const TIn* a00 = at(src, s.p00);
  const TIn* a01 = at(src, s.p01);
  const TIn* a10 = at(src, s.p10);
  const TIn* a11 = at(src, s.p11);
  const bool ld_top = !CACHE || s.p00 != id_top;
  const bool ld_bot = !CACHE || s.p10 != id_bot;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (group_on<FULL>(c0, g, C)) {
      if (ld_top) col[l][0][g] = Io<TIn>::ld(a00 + 128 * g);
      if (ld_bot) col[l][1][g] = Io<TIn>::ld(a10 + 128 * g);
      col[r][0][g] = Io<TIn>::ld(a01 + 128 * g);
      col[r][1][g] = Io<TIn>::ld(a11 + 128 * g);
    }
  }
  id_top = s.p01;
  id_bot = s.p11;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float4 w = f4scale(col[l][0][g], s.w00);
    w = f4fma(col[r][0][g], s.w01, w);
    w = f4fma(col[l][1][g], s.w10, w);
    wv[g] = f4fma(col[r][1][g], s.w11, w);
  }
}

// Pull the RIGHT tap column of a later sample into L1 (the left one is either
// in registers by then or was prefetched as the previous right column).
template <typename TIn, int G, bool FULL>
__device__ __forceinline__ void prefetch_sample(const TIn* __restrict__ src, const WarpSample* sp,
                                                int c0, int C) {
  const uint4 o = *reinterpret_cast<const uint4*>(&sp->p00);
  if (o.x == kNoSample) return;
  const TIn* a01 = at(src, o.y);
  const TIn* a11 = at(src, o.w);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (group_on<FULL>(c0, g, C)) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a01 + 128 * g));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a11 + 128 * g));
    }
  }
}

// host-side helpers shared by the two translation units
int sweep_check(const char* who, int V, int C, int D, int H, int W, int k, int layout);
bool sweep_grid(SweepParams& p, int G, dim3& grid);
int sweep_groups(int C);

}  // namespace mvsd
