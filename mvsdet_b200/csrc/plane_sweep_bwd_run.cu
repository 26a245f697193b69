// Plane-sweep backward, run-merging variant.
//
// ncu on the pixel-per-warp backward (plane_sweep_bwd.cu) shows the kernel bound
// by RED traffic leaving the SM (l1tex2xbar 32 B/clk/SM ~78% busy; L2 atomic unit
// ~50%): every (pixel, plane, neighbour) emits four 1 KB vector REDs.  Here a warp
// walks a horizontal run of kRun pixels of one row (planes in the outer loop) and
// keeps the contribution to the RIGHT tap column pending in registers: when the
// next pixel's LEFT column is the same source pixel (source position advanced by
// exactly one -- the common case between pose-space neighbours) the two
// contributions leave as ONE RED.  Measured: 1.46x fewer RED sectors.
// The per-pixel reference gradient (a sum over planes) lives in shared memory.
#include "plane_sweep.cuh"

namespace mvsd {

#ifndef MVSD_KRUN
#define MVSD_KRUN 8
#endif
#ifndef MVSD_RUNQ_MINB
#define MVSD_RUNQ_MINB 3         // CTAs per SM the lean kernel is compiled for (168 registers)
#endif
constexpr int kRun = MVSD_KRUN;            // pixels per warp run
constexpr int kRunRows = 4;                // rows (= warps) per CTA
constexpr int kRunThreads = kRunRows * 32;
constexpr unsigned kNoTap = 0xfffffffeu;   // "nothing pending"

struct RunCoord {
  int v, y, x0, npix, c0;
};

template <int G>
__device__ __forceinline__ RunCoord run_coord(const SweepParams& p, int warp, int lane) {
  RunCoord c;
  int t = blockIdx.x;
  const int xr = t % p.tiles_x; t /= p.tiles_x;
  const int yt = t % p.tiles_y; t /= p.tiles_y;
  const int slice = t % p.slices;
  c.v = t / p.slices;
  c.y = yt * kRunRows + warp;
  c.x0 = xr * kRun;
  c.npix = min(kRun, p.W - c.x0);
  c.c0 = slice * 128 * G + 4 * lane;
  return c;
}

// lane s -> sample (plane d0 + s / (kRun*k), pixel x0 + (s % (kRun*k)) / k, neighbour s % k)
__device__ __forceinline__ void fill_run_samples(WarpSample* tab, const SweepParams& p,
                                                 const RunCoord& c, int d0, int ppf, int lane) {
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

template <int G, bool FULL>
__device__ __forceinline__ void red_group(float* dst, unsigned off, const float4 (&v)[G], int c0,
                                          int C) {
  float* a = at(dst, off);
#pragma unroll
  for (int g = 0; g < G; ++g)
    if (group_on<FULL>(c0, g, C)) red_add_f32x4(a + 128 * g, v[g]);
}

template <int G, bool FULL>
__device__ __forceinline__ void flush_open(float* dst, unsigned& id, const float4 (&acc)[G], int c0,
                                           int C) {
  if (id != kNoTap) red_group<G, FULL>(dst, id, acc, c0, C);
  id = kNoTap;
}

// One row (top or bottom) of the scatter of one sample: the left tap merges with
// the pending right tap of the previous pixel when it is the same source pixel
// and leaves as one RED; the right tap stays pending.
template <int G, bool FULL>
__device__ __forceinline__ void scatter_side(float* dst, const float4 (&gw)[G], float w_left,
                                             float w_right, unsigned p_left, unsigned p_right,
                                             unsigned& open_id, float4 (&open)[G], int c0, int C) {
  float4 a[G];
  if (open_id == p_left) {
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = f4fma(gw[g], w_left, open[g]);
    red_group<G, FULL>(dst, p_left, a, c0, C);
  } else {
    flush_open<G, FULL>(dst, open_id, open, c0, C);
    if (w_left != 0.f) {
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = f4scale(gw[g], w_left);
      red_group<G, FULL>(dst, p_left, a, c0, C);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) open[g] = f4scale(gw[g], w_right);
  open_id = w_right != 0.f ? p_right : kNoTap;
}

template <typename TIn, typename TG, int KMAX, int G, bool FULL>
__global__ void __launch_bounds__(kRunThreads, 4) sweep_bwd_run_kernel(const SweepParams p) {
  __shared__ WarpSample s_tab[kRunRows][32];
  __shared__ float4 s_gref[kRunRows][kRun][G][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  if (c.y >= p.H) return;
  const int C = p.C, k = p.k, HW = p.H * p.W;
  const TIn* feat = static_cast<const TIn*>(p.feat);
  const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
  const TIn* ref_row = feat + ref_off;
  const TG* g_row = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
  const TIn* nsrc[KMAX];
  float* ndst[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    int n = c.v + p.ref_begin;
    if (j < k) n = __ldg(p.nbr + (size_t)c.v * k + j);
    nsrc[j] = feat + (size_t)n * HW * C + c.c0;
    ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
  }
  const float inv_n = 1.0f / (float)(k + 1);
  const float two_inv_n = 2.0f * inv_n;
  const int spp = kRun * k;
  const int ppf = max(1, 32 / spp);

#pragma unroll
  for (int i = 0; i < kRun; ++i)
#pragma unroll
    for (int g = 0; g < G; ++g) s_gref[warp][i][g][lane] = f4zero();

  for (int d0 = 0; d0 < p.D; d0 += ppf) {
    __syncwarp();
    fill_run_samples(s_tab[warp], p, c, d0, ppf, lane);
    __syncwarp();
    const int dend = min(p.D, d0 + ppf);
    for (int d = d0; d < dend; ++d) {
      float4 open_top[KMAX][G], open_bot[KMAX][G];
      unsigned o_top[KMAX], o_bot[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        o_top[j] = o_bot[j] = kNoTap;
#pragma unroll
        for (int g = 0; g < G; ++g) open_top[j][g] = open_bot[j][g] = f4zero();
      }
      const WarpSample* tab = s_tab[warp] + (d - d0) * spp;
      const TG* g_d = g_row + (size_t)d * HW * C;
#pragma unroll 1
      for (int i = 0; i < c.npix; ++i) {
        float4 gv[G], ref[G], mu[G];
        float4 wv[KMAX][G];
        WarpSample smp[KMAX];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool on = group_on<FULL>(c.c0, g, C);
          gv[g] = on ? Io<TG>::ld_stream(g_d + i * C + 128 * g) : f4zero();
          ref[g] = on ? Io<TIn>::ld(ref_row + i * C + 128 * g) : f4zero();
          mu[g] = ref[g];
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
#pragma unroll
          for (int g = 0; g < G; ++g) wv[j][g] = f4zero();
          if (j >= k) continue;
          smp[j] = tab[i * k + j];
          if (smp[j].p00 == kNoSample) continue;
          gather_taps<TIn, G, FULL>(nsrc[j], smp[j], c.c0, C, wv[j]);
#pragma unroll
          for (int g = 0; g < G; ++g) mu[g] = f4add(mu[g], wv[j][g]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          mu[g] = f4scale(mu[g], inv_n);
          gv[g] = f4scale(gv[g], two_inv_n);
          s_gref[warp][i][g][lane] = f4fma(gv[g], f4sub(ref[g], mu[g]), s_gref[warp][i][g][lane]);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j >= k) continue;
          const WarpSample s = smp[j];
          if (s.p00 == kNoSample) {
            flush_open<G, FULL>(ndst[j], o_top[j], open_top[j], c.c0, C);
            flush_open<G, FULL>(ndst[j], o_bot[j], open_bot[j], c.c0, C);
            continue;
          }
          float4 gw[G];
#pragma unroll
          for (int g = 0; g < G; ++g) gw[g] = f4mul(gv[g], f4sub(wv[j][g], mu[g]));
          scatter_side<G, FULL>(ndst[j], gw, s.w00, s.w01, s.p00, s.p01, o_top[j], open_top[j], c.c0, C);
          scatter_side<G, FULL>(ndst[j], gw, s.w10, s.w11, s.p10, s.p11, o_bot[j], open_bot[j], c.c0, C);
        }
      }
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j >= k) continue;
        flush_open<G, FULL>(ndst[j], o_top[j], open_top[j], c.c0, C);
        flush_open<G, FULL>(ndst[j], o_bot[j], open_bot[j], c.c0, C);
      }
    }
  }
  float* dst = p.g_feat + ref_off;
  for (int i = 0; i < c.npix; ++i) {
#pragma unroll
    for (int g = 0; g < G; ++g)
      if (group_on<FULL>(c.c0, g, C)) red_add_f32x4(dst + i * C + 128 * g, s_gref[warp][i][g][lane]);
  }
}


// ---------------------------------------------------------------------------
// Packed-arithmetic variant (default).  Same run-merging scheme; differences:
//   * fp32 math on f32x2 pairs (FMUL2 / FFMA2 / FADD2): the scalar kernel issues
//     ~390 instructions per pixel-plane and sits at 16 warps/SM;
//   * the upstream-gradient stream (1.18 GB, the only DRAM-sized operand) is
//     pulled into L2 two planes ahead with one cp.async.bulk.prefetch.L2 per
//     warp and plane: ncu showed the scalar kernel waiting on exactly these
//     loads (long-scoreboard 3.6 per issue at ~1 us DRAM latency with ~16 KB in
//     flight per SM), now they are L2 hits and cost no registers;
//   * neighbour base pointers are pinned in registers.
// ---------------------------------------------------------------------------
#ifdef MVSD_EXP_NORED
__constant__ int c_exp_nored;      // experiment: measure the kernel with the neighbour REDs suppressed
#endif
__device__ __forceinline__ void red_add_p4(float* p, P4 v) {
#ifdef MVSD_EXP_NORED
  if (c_exp_nored) return;
#endif
  float a, b, c, d;
  upk2(v.lo, a, b);
  upk2(v.hi, c, d);
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

template <int G, bool FULL>
__device__ __forceinline__ void red_group_p(float* dst, unsigned off, const P4 (&v)[G], int c0,
                                            int C) {
  float* a = at(dst, off);
#pragma unroll
  for (int g = 0; g < G; ++g)
    if (group_on<FULL>(c0, g, C)) red_add_p4(a + 128 * g, v[g]);
}

template <int G, bool FULL>
__device__ __forceinline__ void flush_open_p(float* dst, unsigned& id, const P4 (&acc)[G], int c0,
                                             int C) {
  if (id != kNoTap) red_group_p<G, FULL>(dst, id, acc, c0, C);
  id = kNoTap;
}

template <int G, bool FULL>
__device__ __forceinline__ void scatter_side_p(float* dst, const P4 (&gw)[G], float w_left,
                                               float w_right, unsigned p_left, unsigned p_right,
                                               unsigned& open_id, P4 (&open)[G], int c0, int C) {
  const u64 wl = pk2(w_left, w_left), wr = pk2(w_right, w_right);
  P4 a[G];
  if (open_id == p_left) {
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = p4fma(gw[g], wl, open[g]);
    red_group_p<G, FULL>(dst, p_left, a, c0, C);
  } else {
    flush_open_p<G, FULL>(dst, open_id, open, c0, C);
    if (w_left != 0.f) {
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = p4scale(gw[g], wl);
      red_group_p<G, FULL>(dst, p_left, a, c0, C);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) open[g] = p4scale(gw[g], wr);
  open_id = w_right != 0.f ? p_right : kNoTap;
}

constexpr int kPrefetchPlanes = 2;

// TM = true keeps the per-pixel reference-gradient accumulators in tensor memory
// (kRun * G * 4 columns per CTA) instead of 32 KB of shared memory per CTA: the
// L1 that shared memory was carved out of goes back to the gathered taps.
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int MINB, bool TM>
__global__ void __launch_bounds__(kRunThreads, MINB) sweep_bwd_runp_kernel(const SweepParams p) {
  constexpr int kCols = kRun * G * 4;
  __shared__ WarpSample s_tab[kRunRows][32];
  __shared__ P4 s_gref[TM ? 1 : kRunRows][TM ? 1 : kRun][G][32];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  uint32_t tbase = 0;
  if (TM) tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);
  auto gref_ld = [&](int i, int g) -> P4 {
    return TM ? tmem_ld4(tbase + 4u * (uint32_t)(i * G + g)) : s_gref[TM ? 0 : warp][TM ? 0 : i][g][lane];
  };
  auto gref_st = [&](int i, int g, P4 v) {
    if (TM) tmem_st4(tbase + 4u * (uint32_t)(i * G + g), v);
    else s_gref[TM ? 0 : warp][TM ? 0 : i][g][lane] = v;
  };
  if (c.y < p.H) {
  const int C = p.C, k = p.k, HW = p.H * p.W;
  const TIn* feat = static_cast<const TIn*>(p.feat);
  const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
  const TIn* ref_row = feat + ref_off;
  const size_t plane_stride = (size_t)HW * C;
  const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
  // L2 prefetch of this warp's slice of the run: npix pixels, (128*G or C-c0) channels each;
  // with one channel slice (C <= 128*G) the run is one contiguous chunk.
  const bool one_chunk = p.slices == 1;
  const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                       sizeof(TG)) & ~15u;
  const TG* pf_base = g_d - 4 * lane;            // lane-independent start of the slice
  const bool pf_ok = pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                     ((plane_stride * sizeof(TG)) & 15) == 0;
  auto prefetch_plane = [&](int d) {
    if (!pf_ok || d >= p.D) return;
    const TG* q = pf_base + (size_t)d * plane_stride;
    if (one_chunk) {
      if (lane == 0) prefetch_l2(q, pf_bytes);
    } else if (lane < c.npix) {
      prefetch_l2(q + (size_t)lane * C, pf_bytes);
    }
  };
#pragma unroll
  for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

  const TIn* nsrc[KMAX];
  float* ndst[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    int n = c.v + p.ref_begin;
    if (j < k) n = __ldg(p.nbr + (size_t)c.v * k + j);
    nsrc[j] = feat + (size_t)n * HW * C + c.c0;
    ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
    asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
  }
  const float inv_n = 1.0f / (float)(k + 1);
  const u64 inv_n2 = pk2(inv_n, inv_n);
  const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
  const int spp = kRun * k;
  const int ppf = max(1, 32 / spp);

#pragma unroll
  for (int i = 0; i < kRun; ++i)
#pragma unroll
    for (int g = 0; g < G; ++g) gref_st(i, g, p4zero());
  if (TM) tmem_wait_st();

  for (int d0 = 0; d0 < p.D; d0 += ppf) {
    __syncwarp();
    fill_run_samples(s_tab[warp], p, c, d0, ppf, lane);
    __syncwarp();
    const int dend = min(p.D, d0 + ppf);
    for (int d = d0; d < dend; ++d) {
      prefetch_plane(d + kPrefetchPlanes);
      if (TM) tmem_wait_st();
      P4 open_top[KMAX][G], open_bot[KMAX][G];
      unsigned o_top[KMAX], o_bot[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        o_top[j] = o_bot[j] = kNoTap;
#pragma unroll
        for (int g = 0; g < G; ++g) open_top[j][g] = open_bot[j][g] = p4zero();
      }
      const WarpSample* tab = s_tab[warp] + (d - d0) * spp;
#pragma unroll 1
      for (int i = 0; i < c.npix; ++i) {
        P4 gv[G], ref[G], mu[G];
        P4 wv[KMAX][G];
        WarpSample smp[KMAX];
        typename Raw<TG>::type graw[G];
        typename Raw<TIn>::type rraw[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool on = group_on<FULL>(c.c0, g, C);
          graw[g] = on ? Raw<TG>::ld_stream_na(g_d + i * C + 128 * g) : Raw<TG>::zero();
          rraw[g] = on ? Raw<TIn>::ld(ref_row + i * C + 128 * g) : Raw<TIn>::zero();
        }
        // all loads of the pixel (gradient, reference, both neighbours' taps) go out
        // before the first dependent instruction: one L2 round trip per pixel, not three
        RawTaps<TIn, G> traw[KMAX];
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j >= k) continue;
          smp[j] = tab[i * k + j];
          if (smp[j].p00 != kNoSample) load_taps<TIn, G, FULL>(nsrc[j], smp[j], c.c0, C, traw[j]);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
#pragma unroll
          for (int g = 0; g < G; ++g) wv[j][g] = p4zero();
          if (j >= k || smp[j].p00 == kNoSample) continue;
          blend_taps<TIn, G>(traw[j], smp[j], wv[j]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          ref[g] = p4from(rraw[g]);
          mu[g] = ref[g];
#pragma unroll
          for (int j = 0; j < KMAX; ++j)
            if (j < k) mu[g] = p4add(mu[g], wv[j][g]);
          mu[g] = p4scale(mu[g], inv_n2);
          gv[g] = p4scale(p4from(graw[g]), two_inv_n2);
          gref_st(i, g, p4fma(gv[g], p4sub(ref[g], mu[g]), gref_ld(i, g)));
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j >= k) continue;
          const WarpSample s = smp[j];
          if (s.p00 == kNoSample) {
            flush_open_p<G, FULL>(ndst[j], o_top[j], open_top[j], c.c0, C);
            flush_open_p<G, FULL>(ndst[j], o_bot[j], open_bot[j], c.c0, C);
            continue;
          }
          P4 gw[G];
#pragma unroll
          for (int g = 0; g < G; ++g) gw[g] = p4mul(gv[g], p4sub(wv[j][g], mu[g]));
          scatter_side_p<G, FULL>(ndst[j], gw, s.w00, s.w01, s.p00, s.p01, o_top[j], open_top[j], c.c0, C);
          scatter_side_p<G, FULL>(ndst[j], gw, s.w10, s.w11, s.p10, s.p11, o_bot[j], open_bot[j], c.c0, C);
        }
      }
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j >= k) continue;
        flush_open_p<G, FULL>(ndst[j], o_top[j], open_top[j], c.c0, C);
        flush_open_p<G, FULL>(ndst[j], o_bot[j], open_bot[j], c.c0, C);
      }
      g_d += plane_stride;
    }
  }
  if (TM) tmem_wait_st();
  float* dst = p.g_feat + ref_off;
  for (int i = 0; i < c.npix; ++i) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const P4 acc = gref_ld(i, g);
      if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
    }
  }
  }  // c.y < p.H
  if (TM) tmem_free_cta<kCols>(&s_tmem, warp);
}

// ---------------------------------------------------------------------------
// Lean variant (default since round 1d).  ncu's per-instruction counts on the packed
// kernel above: 335 warp instructions per pixel-plane of which 81 are register moves
// (MOV / IMAD.MOV / CS2R at the joins of "zero, then blend if the sample is valid" and of
// the merge-or-flush branches) and 63 are control flow.  Same algorithm, restructured so
// that nothing is merged at a join:
//   * the pixel body is instantiated per validity pattern of the two samples (a
//     warp-uniform 4-way branch): an absent sample has no loads, no blend, no zeroed
//     vector and leaves the pending taps alone (a later mismatch or the end of the run
//     flushes them);
//   * the scatter keeps one invariant -- "the pending accumulator is zero unless it
//     matches the next left tap" -- so the mismatch arm only issues a RED and zeroes the
//     pending registers in place, and the left tap is always fma(gw, w_left, pending).
// ---------------------------------------------------------------------------
template <int KMAX, int G>
struct RunPending {
  P4 top[KMAX][G], bot[KMAX][G];
  unsigned id_top[KMAX], id_bot[KMAX];
};

template <int G, bool FULL>
__device__ __forceinline__ void side_q(float* dst, const P4 (&gw)[G], float w_left, float w_right,
                                       unsigned p_left, unsigned p_right, unsigned& open_id,
                                       P4 (&open)[G], int c0, int C) {
  const u64 wl = pk2(w_left, w_left), wr = pk2(w_right, w_right);
  P4 a[G];
  if (open_id == p_left) {                 // source x advanced by one: pending + left leave as one RED
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = p4fma(gw[g], wl, open[g]);
    red_group_p<G, FULL>(dst, p_left, a, c0, C);
#pragma unroll
    for (int g = 0; g < G; ++g) open[g] = p4scale(gw[g], wr);
    open_id = w_right != 0.f ? p_right : kNoTap;
  } else if (open_id == p_right && w_right != 0.f) {   // source x did not advance: right joins the pending tap
#pragma unroll
    for (int g = 0; g < G; ++g) open[g] = p4fma(gw[g], wr, open[g]);
    if (w_left != 0.f) {
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = p4scale(gw[g], wl);
      red_group_p<G, FULL>(dst, p_left, a, c0, C);
    }
  } else {
    if (open_id != kNoTap) red_group_p<G, FULL>(dst, open_id, open, c0, C);
    if (w_left != 0.f) {
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = p4scale(gw[g], wl);
      red_group_p<G, FULL>(dst, p_left, a, c0, C);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) open[g] = p4scale(gw[g], wr);
    open_id = w_right != 0.f ? p_right : kNoTap;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}

// side_q for pre-weighted contributions (left / right vectors), used by the receiving
// side of a hand-off.
template <int G, bool FULL>
__device__ __forceinline__ void side_c(float* dst, const P4 (&cl)[G], const P4 (&cr)[G], bool nz_left,
                                       bool nz_right, unsigned p_left, unsigned p_right,
                                       unsigned& open_id, P4 (&open)[G], int c0, int C) {
  P4 a[G];
  if (open_id == p_left) {
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = p4add(cl[g], open[g]);
    red_group_p<G, FULL>(dst, p_left, a, c0, C);
#pragma unroll
    for (int g = 0; g < G; ++g) open[g] = cr[g];
    open_id = nz_right ? p_right : kNoTap;
  } else if (open_id == p_right && nz_right) {
#pragma unroll
    for (int g = 0; g < G; ++g) open[g] = p4add(cr[g], open[g]);
    if (nz_left) red_group_p<G, FULL>(dst, p_left, cl, c0, C);
  } else {
    if (open_id != kNoTap) red_group_p<G, FULL>(dst, open_id, open, c0, C);
    if (nz_left) red_group_p<G, FULL>(dst, p_left, cl, c0, C);
#pragma unroll
    for (int g = 0; g < G; ++g) open[g] = cr[g];
    open_id = nz_right ? p_right : kNoTap;
  }
}

// One stage of a hand-off queue, lane-private columns.  NV = 2: the sender's bottom side already
// weighted (left, right vectors); NV = 1: the sender's un-weighted contribution gw -- the receiver
// applies the sender's two bottom weights, which it reads from the shared sample table anyway:
// half the shared memory per stage, so the queue can be as deep as the run (NSTG = 8) and the
// sending row never waits for the receiving one inside a plane.
template <int G, int NV = 2>
struct HandoffSlot {
  P4 v[NV][G][32];
};

// One neighbour's scatter with the hand-off protocol.  send: give the bottom side to the
// row below; recv: take the row above's bottom side into this row's top side.
template <int G, bool FULL, int NSTG, int NV>
__device__ __forceinline__ void scatter_h(float* dst, const P4 (&gw)[G], const WarpSample& s,
                                          unsigned& id_top, P4 (&top)[G], unsigned& id_bot,
                                          P4 (&bot)[G], bool send, HandoffSlot<G, NV>* out_slot,
                                          unsigned long long* out_full, unsigned long long* out_empty,
                                          unsigned& h_out, bool recv, float up_w10, float up_w11,
                                          HandoffSlot<G, NV>* in_slot, unsigned long long* in_full,
                                          unsigned long long* in_empty, unsigned& h_in, int lane,
                                          int c0, int C) {
  static_assert((NSTG & (NSTG - 1)) == 0, "stage count must be a power of two");
  // bottom side first: an early hand-off unblocks the warp below
  if (send) {
    const unsigned h = h_out++;
    const unsigned stg = h & (unsigned)(NSTG - 1), ph = (h / (unsigned)NSTG) & 1u;
    mbar_wait(out_empty + stg, ph ^ 1u);
    if (NV == 2) {
      const u64 w10 = pk2(s.w10, s.w10), w11 = pk2(s.w11, s.w11);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        out_slot[stg].v[0][g][lane] = p4scale(gw[g], w10);
        out_slot[stg].v[NV - 1][g][lane] = p4scale(gw[g], w11);
      }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g) out_slot[stg].v[0][g][lane] = gw[g];
    }
    mbar_arrive(out_full + stg);
  } else {
    side_q<G, FULL>(dst, gw, s.w10, s.w11, s.p10, s.p11, id_bot, bot, c0, C);
  }
  if (recv) {
    const unsigned h = h_in++;
    const unsigned stg = h & (unsigned)(NSTG - 1), ph = (h / (unsigned)NSTG) & 1u;
    mbar_wait(in_full + stg, ph);
    P4 cl[G], cr[G];
    if (NV == 2) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        cl[g] = in_slot[stg].v[0][g][lane];
        cr[g] = in_slot[stg].v[NV - 1][g][lane];
      }
    } else {
      const u64 u10 = pk2(up_w10, up_w10), u11 = pk2(up_w11, up_w11);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 gu = in_slot[stg].v[0][g][lane];
        cl[g] = p4scale(gu, u10);
        cr[g] = p4scale(gu, u11);
      }
    }
    mbar_arrive(in_empty + stg);
    const u64 w00 = pk2(s.w00, s.w00), w01 = pk2(s.w01, s.w01);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cl[g] = p4fma(gw[g], w00, cl[g]);
      cr[g] = p4fma(gw[g], w01, cr[g]);
    }
    side_c<G, FULL>(dst, cl, cr, s.w00 != 0.f || up_w10 != 0.f, s.w01 != 0.f || up_w11 != 0.f, s.p00,
                    s.p01, id_top, top, c0, C);
  } else {
    side_q<G, FULL>(dst, gw, s.w00, s.w01, s.p00, s.p01, id_top, top, c0, C);
  }
}

template <int KMAX, int G, int NSTG = 2, int NV = 2>
struct HandoffCtx {                  // per-pixel hand-off decisions + the CTA's slots and barriers
  static constexpr int kStages = NSTG, kVec = NV;
  bool send[KMAX], recv[KMAX];
  float up_w10[KMAX], up_w11[KMAX];
  HandoffSlot<G, NV> (*out_slot)[NSTG];     // [KMAX][NSTG] of the boundary below this row
  unsigned long long (*out_full)[NSTG], (*out_empty)[NSTG];
  HandoffSlot<G, NV> (*in_slot)[NSTG];      // boundary above this row
  unsigned long long (*in_full)[NSTG], (*in_empty)[NSTG];
  unsigned h_out[KMAX], h_in[KMAX];
  int lane;
};

template <typename TIn, typename TG, int KMAX, int G, bool FULL, bool V0, bool V1, bool HO = false,
          typename HOC = HandoffCtx<KMAX, G>>
__device__ __forceinline__ void pixel_q(RunPending<KMAX, G>& pend, const WarpSample& s0,
                                        const WarpSample& s1, const TG* __restrict__ gp,
                                        const TIn* __restrict__ rp, const TIn* const (&nsrc)[KMAX],
                                        float* const (&ndst)[KMAX], uint32_t taddr, u64 inv_n2,
                                        u64 two_inv_n2, int c0, int C, HOC* ho = nullptr) {
  typename Raw<TG>::type graw[G];
  typename Raw<TIn>::type rraw[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const bool on = group_on<FULL>(c0, g, C);
    graw[g] = on ? Raw<TG>::ld_stream_na(gp + 128 * g) : Raw<TG>::zero();
    rraw[g] = on ? Raw<TIn>::ld(rp + 128 * g) : Raw<TIn>::zero();
  }
  RawTaps<TIn, G> t0, t1;
  if (V0) load_taps<TIn, G, FULL>(nsrc[0], s0, c0, C, t0);
  if (V1) load_taps<TIn, G, FULL>(nsrc[KMAX - 1], s1, c0, C, t1);
  P4 w0[G], w1[G], gw0[G], gw1[G];
  if (V0) blend_taps<TIn, G>(t0, s0, w0);
  if (V1) blend_taps<TIn, G>(t1, s1, w1);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const P4 ref = p4from(rraw[g]);
    P4 mu = ref;
    if (V0) mu = p4add(mu, w0[g]);
    if (V1) mu = p4add(mu, w1[g]);
    mu = p4scale(mu, inv_n2);
    const P4 gv = p4scale(p4from(graw[g]), two_inv_n2);
    const uint32_t ta = taddr + 4u * (uint32_t)g;
    tmem_st4(ta, p4fma(gv, p4sub(ref, mu), tmem_ld4(ta)));
    if (V0) gw0[g] = p4mul(gv, p4sub(w0[g], mu));
    if (V1) gw1[g] = p4mul(gv, p4sub(w1[g], mu));
  }
  if (V0) {
    if (HO) {
      scatter_h<G, FULL, HOC::kStages, HOC::kVec>(ndst[0], gw0, s0, pend.id_top[0], pend.top[0], pend.id_bot[0], pend.bot[0],
                         ho->send[0], ho->out_slot[0], ho->out_full[0], ho->out_empty[0], ho->h_out[0],
                         ho->recv[0], ho->up_w10[0], ho->up_w11[0], ho->in_slot[0], ho->in_full[0],
                         ho->in_empty[0], ho->h_in[0], ho->lane, c0, C);
    } else {
      side_q<G, FULL>(ndst[0], gw0, s0.w00, s0.w01, s0.p00, s0.p01, pend.id_top[0], pend.top[0], c0, C);
      side_q<G, FULL>(ndst[0], gw0, s0.w10, s0.w11, s0.p10, s0.p11, pend.id_bot[0], pend.bot[0], c0, C);
    }
  }
  if (V1) {
    constexpr int J = KMAX - 1;
    if (HO) {
      scatter_h<G, FULL, HOC::kStages, HOC::kVec>(ndst[J], gw1, s1, pend.id_top[J], pend.top[J], pend.id_bot[J], pend.bot[J],
                         ho->send[J], ho->out_slot[J], ho->out_full[J], ho->out_empty[J], ho->h_out[J],
                         ho->recv[J], ho->up_w10[J], ho->up_w11[J], ho->in_slot[J], ho->in_full[J],
                         ho->in_empty[J], ho->h_in[J], ho->lane, c0, C);
    } else {
      side_q<G, FULL>(ndst[J], gw1, s1.w00, s1.w01, s1.p00, s1.p01, pend.id_top[J], pend.top[J], c0, C);
      side_q<G, FULL>(ndst[J], gw1, s1.w10, s1.w11, s1.p10, s1.p11, pend.id_bot[J], pend.bot[J], c0, C);
    }
  }
}

// requires p.k == KMAX (the launcher instantiates KMAX = k for k in {1, 2})
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int MINB>
__global__ void __launch_bounds__(kRunThreads, MINB) sweep_bwd_runq_kernel(const SweepParams p) {
  constexpr int kCols = kRun * G * 4;
  __shared__ WarpSample s_tab[kRunRows][32];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  const uint32_t tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);
  if (c.y < p.H) {
    const int C = p.C, HW = p.H * p.W;
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const TIn* ref_row = feat + ref_off;
    const size_t plane_stride = (size_t)HW * C;
    const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const bool one_chunk = p.slices == 1;
    const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                         sizeof(TG)) & ~15u;
    const TG* pf_base = g_d - 4 * lane;
    const bool pf_ok = pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                       ((plane_stride * sizeof(TG)) & 15) == 0;
    auto prefetch_plane = [&](int d) {
      if (!pf_ok || d >= p.D) return;
      const TG* q = pf_base + (size_t)d * plane_stride;
      if (one_chunk) {
        if (lane == 0) prefetch_l2(q, pf_bytes);
      } else if (lane < c.npix) {
        prefetch_l2(q + (size_t)lane * C, pf_bytes);
      }
    };
#pragma unroll
    for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int n = __ldg(p.nbr + (size_t)c.v * KMAX + j);
      nsrc[j] = feat + (size_t)n * HW * C + c.c0;
      ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(KMAX + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
    constexpr int spp = kRun * KMAX;
    constexpr int ppf = 32 / spp > 0 ? 32 / spp : 1;

#pragma unroll
    for (int q = 0; q < kRun * G; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();

    for (int d0 = 0; d0 < p.D; d0 += ppf) {
      __syncwarp();
      fill_run_samples(s_tab[warp], p, c, d0, ppf, lane);
      __syncwarp();
      const int dend = min(p.D, d0 + ppf);
      for (int d = d0; d < dend; ++d) {
        prefetch_plane(d + kPrefetchPlanes);
        tmem_wait_st();
        RunPending<KMAX, G> pend;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          pend.id_top[j] = pend.id_bot[j] = kNoTap;
#pragma unroll
          for (int g = 0; g < G; ++g) pend.top[j][g] = pend.bot[j][g] = p4zero();
        }
        const WarpSample* tab = s_tab[warp] + (d - d0) * spp;
#pragma unroll 1
        for (int i = 0; i < c.npix; ++i) {
          const WarpSample s0 = tab[i * KMAX];
          const WarpSample s1 = tab[i * KMAX + (KMAX - 1)];
          const bool v0 = s0.p00 != kNoSample;
          const bool v1 = KMAX == 2 && s1.p00 != kNoSample;
          const TG* gp = g_d + i * C;
          const TIn* rp = ref_row + i * C;
          const uint32_t ta = tbase + 4u * (uint32_t)(i * G);
          if (v0 && v1)
            pixel_q<TIn, TG, KMAX, G, FULL, true, true>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
          else if (v0)
            pixel_q<TIn, TG, KMAX, G, FULL, true, false>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
          else if (v1)
            pixel_q<TIn, TG, KMAX, G, FULL, false, true>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
          else
            pixel_q<TIn, TG, KMAX, G, FULL, false, false>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          flush_open_p<G, FULL>(ndst[j], pend.id_top[j], pend.top[j], c.c0, C);
          flush_open_p<G, FULL>(ndst[j], pend.id_bot[j], pend.bot[j], c.c0, C);
        }
        g_d += plane_stride;
      }
    }
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int i = 0; i < c.npix; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(i * G + g));
        if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
      }
    }
  }
  tmem_free_cta<kCols>(&s_tmem, warp);
}

// ---------------------------------------------------------------------------
// Software-pipelined lean kernel (tuning key 5 = 14; 5 = 7 selects the un-pipelined sweep_bwd_runq
// above).  ncu's per-warp picture of the lean kernel:
// ~2260 cycles per pixel-plane = one L2 round trip for the 20 loads of the pixel (all issued
// together, ~1000+ cycles under the RED traffic) followed by ~240 dependent-ish instructions at
// ~4 cycles each, with only 3 warps per scheduler to overlap the two.  An L1 prefetch of the next
// pixel hides the round trip but doubles the L1 tag traffic and is slower (measured twice).  Here
// the loads of the NEXT pixel are issued into the raw-load registers as soon as the blend has
// consumed the current ones, i.e. before the variance algebra, the tensor-memory update and the
// scatter: the same registers, no extra L1 traffic, and the round trip overlaps ~60 % of the
// pixel's instructions.  The pipeline runs across the planes of one sample-table fill.  (With the
// un-slimmed hand-off kernel the same pipelining needed 168 registers + 80 bytes of spills: slower.)
// ---------------------------------------------------------------------------
template <typename TIn, typename TG, int G>
struct PixelRaw {
  typename Raw<TG>::type g[G];
  typename Raw<TIn>::type r[G];
  RawTaps<TIn, G> t0, t1;
};

template <typename TIn, typename TG, int KMAX, int G, bool FULL>
__device__ __forceinline__ void issue_pixel_loads(PixelRaw<TIn, TG, G>& raw, const WarpSample* smp,
                                                  const TG* __restrict__ gp, const TIn* __restrict__ rp,
                                                  const TIn* const (&nsrc)[KMAX], int c0, int C) {
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const bool on = group_on<FULL>(c0, g, C);
    raw.g[g] = on ? Raw<TG>::ld_stream_na(gp + 128 * g) : Raw<TG>::zero();
    raw.r[g] = on ? Raw<TIn>::ld(rp + 128 * g) : Raw<TIn>::zero();
  }
  // only the four tap offsets are needed to issue the loads: the second 16 bytes of a sample
  WarpSample a;
  const uint4 q0 = reinterpret_cast<const uint4*>(smp)[1];
  a.p00 = q0.x; a.p01 = q0.y; a.p10 = q0.z; a.p11 = q0.w;
  if (q0.x != kNoSample) load_taps<TIn, G, FULL>(nsrc[0], a, c0, C, raw.t0);
  if (KMAX == 2) {
    const uint4 q1 = reinterpret_cast<const uint4*>(smp + (KMAX - 1))[1];
    a.p00 = q1.x; a.p01 = q1.y; a.p10 = q1.z; a.p11 = q1.w;
    if (q1.x != kNoSample) load_taps<TIn, G, FULL>(nsrc[KMAX - 1], a, c0, C, raw.t1);
  }
}

template <typename TIn, typename TG, int KMAX, int G, bool FULL, bool V0, bool V1>
__device__ __forceinline__ void pixel_q2(RunPending<KMAX, G>& pend, PixelRaw<TIn, TG, G>& raw,
                                         const WarpSample& s0, const WarpSample& s1, bool has_next,
                                         const WarpSample* smp_next, const TG* __restrict__ gp_next,
                                         const TIn* __restrict__ rp_next, const TIn* const (&nsrc)[KMAX],
                                         float* const (&ndst)[KMAX], uint32_t taddr, u64 inv_n2,
                                         u64 two_inv_n2, int c0, int C) {
  P4 w0[G], w1[G], gw0[G], gw1[G], ref[G], gv[G];
  if (V0) blend_taps<TIn, G>(raw.t0, s0, w0);
  if (V1) blend_taps<TIn, G>(raw.t1, s1, w1);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ref[g] = p4from(raw.r[g]);
    gv[g] = p4scale(p4from(raw.g[g]), two_inv_n2);
  }
  // the raw registers are free: the next pixel's round trip starts here
  if (has_next) issue_pixel_loads<TIn, TG, KMAX, G, FULL>(raw, smp_next, gp_next, rp_next, nsrc, c0, C);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    P4 mu = ref[g];
    if (V0) mu = p4add(mu, w0[g]);
    if (V1) mu = p4add(mu, w1[g]);
    mu = p4scale(mu, inv_n2);
    const uint32_t ta = taddr + 4u * (uint32_t)g;
    tmem_st4(ta, p4fma(gv[g], p4sub(ref[g], mu), tmem_ld4(ta)));
    if (V0) gw0[g] = p4mul(gv[g], p4sub(w0[g], mu));
    if (V1) gw1[g] = p4mul(gv[g], p4sub(w1[g], mu));
  }
  if (V0) {
    side_q<G, FULL>(ndst[0], gw0, s0.w00, s0.w01, s0.p00, s0.p01, pend.id_top[0], pend.top[0], c0, C);
    side_q<G, FULL>(ndst[0], gw0, s0.w10, s0.w11, s0.p10, s0.p11, pend.id_bot[0], pend.bot[0], c0, C);
  }
  if (V1) {
    constexpr int J = KMAX - 1;
    side_q<G, FULL>(ndst[J], gw1, s1.w00, s1.w01, s1.p00, s1.p01, pend.id_top[J], pend.top[J], c0, C);
    side_q<G, FULL>(ndst[J], gw1, s1.w10, s1.w11, s1.p10, s1.p11, pend.id_bot[J], pend.bot[J], c0, C);
  }
}

// requires p.k == KMAX
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int MINB>
__global__ void __launch_bounds__(kRunThreads, MINB) sweep_bwd_runq2_kernel(const SweepParams p) {
  constexpr int kCols = kRun * G * 4;
  __shared__ WarpSample s_tab[kRunRows][32];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  const uint32_t tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);
  if (c.y < p.H) {
    const int C = p.C, HW = p.H * p.W;
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const TIn* ref_row = feat + ref_off;
    const size_t plane_stride = (size_t)HW * C;
    const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const bool one_chunk = p.slices == 1;
    const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                         sizeof(TG)) & ~15u;
    const TG* pf_base = g_d - 4 * lane;
    const bool pf_ok = pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                       ((plane_stride * sizeof(TG)) & 15) == 0;
    auto prefetch_plane = [&](int d) {
      if (!pf_ok || d >= p.D) return;
      const TG* q = pf_base + (size_t)d * plane_stride;
      if (one_chunk) {
        if (lane == 0) prefetch_l2(q, pf_bytes);
      } else if (lane < c.npix) {
        prefetch_l2(q + (size_t)lane * C, pf_bytes);
      }
    };
#pragma unroll
    for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int n = __ldg(p.nbr + (size_t)c.v * KMAX + j);
      nsrc[j] = feat + (size_t)n * HW * C + c.c0;
      ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(KMAX + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
    constexpr int spp = kRun * KMAX;
    constexpr int ppf = 32 / spp > 0 ? 32 / spp : 1;

#pragma unroll
    for (int q = 0; q < kRun * G; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();

    PixelRaw<TIn, TG, G> raw;
    for (int d0 = 0; d0 < p.D; d0 += ppf) {
      __syncwarp();
      fill_run_samples(s_tab[warp], p, c, d0, ppf, lane);
      __syncwarp();
      const int dend = min(p.D, d0 + ppf);
      // pipeline prologue: first pixel of the first plane of this table fill
      issue_pixel_loads<TIn, TG, KMAX, G, FULL>(raw, s_tab[warp], g_d, ref_row, nsrc, c.c0, C);
      for (int d = d0; d < dend; ++d) {
        prefetch_plane(d + kPrefetchPlanes);
        tmem_wait_st();
        RunPending<KMAX, G> pend;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          pend.id_top[j] = pend.id_bot[j] = kNoTap;
#pragma unroll
          for (int g = 0; g < G; ++g) pend.top[j][g] = pend.bot[j][g] = p4zero();
        }
        const WarpSample* tab = s_tab[warp] + (d - d0) * spp;
        const bool more_planes = d + 1 < dend;
#pragma unroll 1
        for (int i = 0; i < c.npix; ++i) {
          const WarpSample s0 = tab[i * KMAX];
          const WarpSample s1 = tab[i * KMAX + (KMAX - 1)];
          const bool v0 = s0.p00 != kNoSample;
          const bool v1 = KMAX == 2 && s1.p00 != kNoSample;
          const bool in_run = i + 1 < c.npix;
          const bool has_next = in_run || more_planes;
          const WarpSample* smp_next = in_run ? tab + (i + 1) * KMAX : tab + spp;
          const TG* gp_next = in_run ? g_d + (i + 1) * C : g_d + plane_stride;
          const TIn* rp_next = in_run ? ref_row + (i + 1) * C : ref_row;
          const uint32_t ta = tbase + 4u * (uint32_t)(i * G);
          if (v0 && v1)
            pixel_q2<TIn, TG, KMAX, G, FULL, true, true>(pend, raw, s0, s1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
          else if (v0)
            pixel_q2<TIn, TG, KMAX, G, FULL, true, false>(pend, raw, s0, s1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
          else if (v1)
            pixel_q2<TIn, TG, KMAX, G, FULL, false, true>(pend, raw, s0, s1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
          else
            pixel_q2<TIn, TG, KMAX, G, FULL, false, false>(pend, raw, s0, s1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          flush_open_p<G, FULL>(ndst[j], pend.id_top[j], pend.top[j], c.c0, C);
          flush_open_p<G, FULL>(ndst[j], pend.id_bot[j], pend.bot[j], c.c0, C);
        }
        g_d += plane_stride;
      }
    }
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int i = 0; i < c.npix; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(i * G + g));
        if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
      }
    }
  }
  tmem_free_cta<kCols>(&s_tmem, warp);
}

// ---------------------------------------------------------------------------
// Row-handoff variant (tuning key 5 = 8).  The lean kernel is bound by the RED stream
// (3.6 GB against a 5.7 TB/s ceiling): a further cut has to merge across image rows.
// The four warps of a CTA own four consecutive rows of the same 8-pixel run; the bottom
// taps of row r and the top taps of row r+1 are usually the same two source pixels.
// When they are (decided identically by both warps from the shared sample tables), warp
// r does not scatter its bottom contribution: it hands the two weighted vectors to warp
// r+1 through a double-buffered shared-memory slot (mbarrier full/empty pair, 32
// arrivals each), and warp r+1 adds them to its own top contribution before its
// merge-or-flush step.  Replay (tools/red_merge_sim.py): 3.5 -> 2.7 GB of RED payload.
// ---------------------------------------------------------------------------
// requires p.k == KMAX
template <int KMAX, int G, int NSTG, int NV>
constexpr size_t runh_slot_bytes() { return sizeof(HandoffSlot<G, NV>) * (kRunRows - 1) * KMAX * NSTG; }

// NSTG stages of NV vectors per (row boundary, neighbour): <2, 2> is what is built (tuning 5 = 8).  Deeper
// queues of un-weighted contributions (<4, 1>, <8, 1>) were measured in round 1f: 0.862 ms against 0.889 ms,
// no difference between 4 and 8 stages.  The slots live in dynamic shared memory.
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int MINB, int NSTG = 2, int NV = 2>
__global__ void __launch_bounds__(kRunThreads, MINB) sweep_bwd_runh_kernel(const SweepParams p) {
  constexpr int kCols = kRun * G * 4;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  typedef HandoffSlot<G, NV> (*SlotArr)[KMAX][NSTG];
  SlotArr s_slot = reinterpret_cast<SlotArr>(s_dyn);          // [kRunRows - 1][KMAX][NSTG]
  __shared__ WarpSample s_tab[kRunRows][32];
  __shared__ __align__(8) unsigned long long s_full[kRunRows - 1][KMAX][NSTG];
  __shared__ __align__(8) unsigned long long s_empty[kRunRows - 1][KMAX][NSTG];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  if (threadIdx.x < (kRunRows - 1) * KMAX * NSTG) {
    mbar_init(&s_full[0][0][0] + threadIdx.x, 32);
    mbar_init(&s_empty[0][0][0] + threadIdx.x, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);     // contains the CTA barriers
  const bool active = c.y < p.H;
  const int C = p.C, HW = p.H * p.W;
  const TIn* feat = static_cast<const TIn*>(p.feat);
  const int yy = active ? c.y : 0;
  const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)yy * p.W + c.x0) * C + c.c0;
  const TIn* ref_row = feat + ref_off;
  const size_t plane_stride = (size_t)HW * C;
  const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)yy * p.W + c.x0) * C + c.c0;
  const bool one_chunk = p.slices == 1;
  const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                       sizeof(TG)) & ~15u;
  const TG* pf_base = g_d - 4 * lane;
  const bool pf_ok = active && pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                     ((plane_stride * sizeof(TG)) & 15) == 0;
  auto prefetch_plane = [&](int d) {
    if (!pf_ok || d >= p.D) return;
    const TG* q = pf_base + (size_t)d * plane_stride;
    if (one_chunk) {
      if (lane == 0) prefetch_l2(q, pf_bytes);
    } else if (lane < c.npix) {
      prefetch_l2(q + (size_t)lane * C, pf_bytes);
    }
  };
#pragma unroll
  for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

  const TIn* nsrc[KMAX];
  float* ndst[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    const int n = __ldg(p.nbr + (size_t)c.v * KMAX + j);
    nsrc[j] = feat + (size_t)n * HW * C + c.c0;
    ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
    asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
  }
  const float inv_n = 1.0f / (float)(KMAX + 1);
  const u64 inv_n2 = pk2(inv_n, inv_n);
  const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
  constexpr int spp = kRun * KMAX;
  constexpr int ppf = 32 / spp > 0 ? 32 / spp : 1;
  const int wdn = min(warp + 1, kRunRows - 1), wup = max(warp - 1, 0);
  const int bo = min(warp, kRunRows - 2), bi = max(warp - 1, 0);     // boundary below / above this row
  typedef HandoffCtx<KMAX, G, NSTG, NV> HOC;
  HOC ho;
  ho.out_slot = s_slot[bo]; ho.out_full = s_full[bo]; ho.out_empty = s_empty[bo];
  ho.in_slot = s_slot[bi]; ho.in_full = s_full[bi]; ho.in_empty = s_empty[bi];
  ho.lane = lane;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) ho.h_out[j] = ho.h_in[j] = 0u;

  if (active) {
#pragma unroll
    for (int q = 0; q < kRun * G; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();
  }

  for (int d0 = 0; d0 < p.D; d0 += ppf) {
    __syncthreads();                       // every warp is done with the previous tables
    if (active) {
      fill_run_samples(s_tab[warp], p, c, d0, ppf, lane);
    } else {
      WarpSample none;
      none.w00 = none.w01 = none.w10 = none.w11 = 0.f;
      none.p00 = none.p01 = none.p10 = none.p11 = kNoSample;
      s_tab[warp][lane] = none;
    }
    __syncthreads();
    if (!active) continue;
    const int dend = min(p.D, d0 + ppf);
    for (int d = d0; d < dend; ++d) {
      prefetch_plane(d + kPrefetchPlanes);
      tmem_wait_st();
      RunPending<KMAX, G> pend;
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        pend.id_top[j] = pend.id_bot[j] = kNoTap;
#pragma unroll
        for (int g = 0; g < G; ++g) pend.top[j][g] = pend.bot[j][g] = p4zero();
      }
      const int toff = (d - d0) * spp;
      const WarpSample* tab = s_tab[warp] + toff;
      const WarpSample* tab_dn = s_tab[wdn] + toff;
      const WarpSample* tab_up = s_tab[wup] + toff;
#pragma unroll 1
      for (int i = 0; i < c.npix; ++i) {
        const WarpSample s0 = tab[i * KMAX];
        const WarpSample s1 = tab[i * KMAX + (KMAX - 1)];
        const bool v0 = s0.p00 != kNoSample;
        const bool v1 = KMAX == 2 && s1.p00 != kNoSample;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          const WarpSample& sj = j == 0 ? s0 : s1;
          const bool vj = j == 0 ? v0 : v1;
          ho.send[j] = ho.recv[j] = false;
          ho.up_w10[j] = ho.up_w11[j] = 0.f;
          if (vj) {
            if (warp + 1 < kRunRows) {
              const WarpSample& dn = tab_dn[i * KMAX + j];
              ho.send[j] = dn.p00 != kNoSample && dn.p00 == sj.p10 && dn.p01 == sj.p11;
            }
            if (warp > 0) {
              const WarpSample& up = tab_up[i * KMAX + j];
              ho.recv[j] = up.p00 != kNoSample && up.p10 == sj.p00 && up.p11 == sj.p01;
              ho.up_w10[j] = up.w10;
              ho.up_w11[j] = up.w11;
            }
          }
        }
        const TG* gp = g_d + i * C;
        const TIn* rp = ref_row + i * C;
        const uint32_t ta = tbase + 4u * (uint32_t)(i * G);
        if (v0 && v1)
          pixel_q<TIn, TG, KMAX, G, FULL, true, true, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
        else if (v0)
          pixel_q<TIn, TG, KMAX, G, FULL, true, false, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
        else if (v1)
          pixel_q<TIn, TG, KMAX, G, FULL, false, true, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
        else
          pixel_q<TIn, TG, KMAX, G, FULL, false, false, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
      }
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        flush_open_p<G, FULL>(ndst[j], pend.id_top[j], pend.top[j], c.c0, C);
        flush_open_p<G, FULL>(ndst[j], pend.id_bot[j], pend.bot[j], c.c0, C);
      }
      g_d += plane_stride;
    }
  }
  if (active) {
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int i = 0; i < c.npix; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(i * G + g));
        if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
      }
    }
  }
  tmem_free_cta<kCols>(&s_tmem, warp);
}

// ---------------------------------------------------------------------------
// Row hand-off with the decisions taken at table-fill time (tuning key 5 = 11).  ncu on sweep_bwd_runh: 366 M warp instructions against 274 M for the lean
// kernel and a CTA barrier every two planes -- the per-pixel send / receive tests read the
// sample tables of the rows above and below (six extra 16-byte shared loads and ~30 ALU
// instructions per pixel), which is also what forces the four warps to refill their tables
// in lock step.  Here the lane that computes a sample also computes the samples of the two
// adjacent rows (same function, same inputs: the same bits the other warp gets) and stores
// the outcome as four flag bits next to the sample; a warp then reads only its own table,
// the per-pixel decision is a byte load, and the only coupling left between the warps of a
// CTA is the hand-off queue itself.
// ---------------------------------------------------------------------------
constexpr unsigned kHoSend = 1u, kHoRecv = 2u, kHoNzLeft = 4u, kHoNzRight = 8u;

__device__ __forceinline__ void fill_run_samples_ho(WarpSample* tab, unsigned char* flg,
                                                    const SweepParams& p, const RunCoord& c, int d0,
                                                    int ppf, int lane, bool has_up, bool has_dn) {
  const int k = p.k, spp = kRun * k;
  if (lane < ppf * spp) {
    const int dd = lane / spp, rem = lane - dd * spp;
    const int i = rem / k, j = rem - i * k;
    const int d = d0 + dd;
    WarpSample s;
    s.w00 = s.w01 = s.w10 = s.w11 = 0.f;
    s.p00 = s.p01 = s.p10 = s.p11 = kNoSample;
    unsigned f = 0u;
    if (d < p.D && i < c.npix) {
      const float* m = p.hom + ((size_t)c.v * k + j) * 12;
      float mm[12];
#pragma unroll
      for (int t = 0; t < 12; ++t) mm[t] = __ldg(m + t);
      const float depth = __ldg(p.depth + (size_t)c.v * p.D + d);
      const float x = (float)(c.x0 + i);
      s = make_warp_sample(mm, x, (float)c.y, depth, p.H, p.W, p.C);
      if (s.p00 != kNoSample) {
        if (has_dn) {
          const WarpSample dn = make_warp_sample(mm, x, (float)(c.y + 1), depth, p.H, p.W, p.C);
          if (dn.p00 != kNoSample && dn.p00 == s.p10 && dn.p01 == s.p11) f |= kHoSend;
        }
        if (has_up) {
          const WarpSample up = make_warp_sample(mm, x, (float)(c.y - 1), depth, p.H, p.W, p.C);
          if (up.p00 != kNoSample && up.p10 == s.p00 && up.p11 == s.p01) {
            f |= kHoRecv;
            if (s.w00 != 0.f || up.w10 != 0.f) f |= kHoNzLeft;
            if (s.w01 != 0.f || up.w11 != 0.f) f |= kHoNzRight;
          }
        }
      }
    }
    tab[lane] = s;
    flg[lane] = (unsigned char)f;
  }
}

// requires p.k == KMAX
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int MINB, int NSTG>
__global__ void __launch_bounds__(kRunThreads, MINB) sweep_bwd_runf_kernel(const SweepParams p) {
  constexpr int kCols = kRun * G * 4;
  constexpr int NV = 2;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  typedef HandoffSlot<G, NV> (*SlotArr)[KMAX][NSTG];
  SlotArr s_slot = reinterpret_cast<SlotArr>(s_dyn);          // [kRunRows - 1][KMAX][NSTG]
  __shared__ WarpSample s_tab[kRunRows][32];
  __shared__ unsigned char s_flg[kRunRows][32];
  __shared__ __align__(8) unsigned long long s_full[kRunRows - 1][KMAX][NSTG];
  __shared__ __align__(8) unsigned long long s_empty[kRunRows - 1][KMAX][NSTG];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  if (threadIdx.x < (kRunRows - 1) * KMAX * NSTG) {
    mbar_init(&s_full[0][0][0] + threadIdx.x, 32);
    mbar_init(&s_empty[0][0][0] + threadIdx.x, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);     // contains the CTA barriers
  if (c.y < p.H) {
    const int C = p.C, HW = p.H * p.W;
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const TIn* ref_row = feat + ref_off;
    const size_t plane_stride = (size_t)HW * C;
    const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const bool one_chunk = p.slices == 1;
    const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                         sizeof(TG)) & ~15u;
    const TG* pf_base = g_d - 4 * lane;
    const bool pf_ok = pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                       ((plane_stride * sizeof(TG)) & 15) == 0;
    auto prefetch_plane = [&](int d) {
      if (!pf_ok || d >= p.D) return;
      const TG* q = pf_base + (size_t)d * plane_stride;
      if (one_chunk) {
        if (lane == 0) prefetch_l2(q, pf_bytes);
      } else if (lane < c.npix) {
        prefetch_l2(q + (size_t)lane * C, pf_bytes);
      }
    };
#pragma unroll
    for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int n = __ldg(p.nbr + (size_t)c.v * KMAX + j);
      nsrc[j] = feat + (size_t)n * HW * C + c.c0;
      ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(KMAX + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
    constexpr int spp = kRun * KMAX;
    constexpr int ppf = 32 / spp > 0 ? 32 / spp : 1;
    const bool has_up = warp > 0;
    const bool has_dn = warp + 1 < kRunRows && c.y + 1 < p.H;
    const int bo = min(warp, kRunRows - 2), bi = max(warp - 1, 0);     // boundary below / above this row
    typedef HandoffCtx<KMAX, G, NSTG, NV> HOC;
    HOC ho;
    ho.out_slot = s_slot[bo]; ho.out_full = s_full[bo]; ho.out_empty = s_empty[bo];
    ho.in_slot = s_slot[bi]; ho.in_full = s_full[bi]; ho.in_empty = s_empty[bi];
    ho.lane = lane;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) ho.h_out[j] = ho.h_in[j] = 0u;

#pragma unroll
    for (int q = 0; q < kRun * G; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();

    for (int d0 = 0; d0 < p.D; d0 += ppf) {
      __syncwarp();
      fill_run_samples_ho(s_tab[warp], s_flg[warp], p, c, d0, ppf, lane, has_up, has_dn);
      __syncwarp();
      const int dend = min(p.D, d0 + ppf);
      for (int d = d0; d < dend; ++d) {
        prefetch_plane(d + kPrefetchPlanes);
        tmem_wait_st();
        RunPending<KMAX, G> pend;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          pend.id_top[j] = pend.id_bot[j] = kNoTap;
#pragma unroll
          for (int g = 0; g < G; ++g) pend.top[j][g] = pend.bot[j][g] = p4zero();
        }
        const int toff = (d - d0) * spp;
        const WarpSample* tab = s_tab[warp] + toff;
        const unsigned char* flg = s_flg[warp] + toff;
#pragma unroll 1
        for (int i = 0; i < c.npix; ++i) {
          const WarpSample s0 = tab[i * KMAX];
          const WarpSample s1 = tab[i * KMAX + (KMAX - 1)];
          const bool v0 = s0.p00 != kNoSample;
          const bool v1 = KMAX == 2 && s1.p00 != kNoSample;
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            const unsigned f = flg[i * KMAX + j];
            ho.send[j] = (f & kHoSend) != 0u;
            ho.recv[j] = (f & kHoRecv) != 0u;
            // scatter_h only tests these against zero (NV == 2: the sender pre-weights)
            ho.up_w10[j] = (f & kHoNzLeft) ? 1.f : 0.f;
            ho.up_w11[j] = (f & kHoNzRight) ? 1.f : 0.f;
          }
          const TG* gp = g_d + i * C;
          const TIn* rp = ref_row + i * C;
          const uint32_t ta = tbase + 4u * (uint32_t)(i * G);
          if (v0 && v1)
            pixel_q<TIn, TG, KMAX, G, FULL, true, true, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
          else if (v0)
            pixel_q<TIn, TG, KMAX, G, FULL, true, false, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
          else if (v1)
            pixel_q<TIn, TG, KMAX, G, FULL, false, true, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
          else
            pixel_q<TIn, TG, KMAX, G, FULL, false, false, true, HOC>(pend, s0, s1, gp, rp, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, &ho);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          flush_open_p<G, FULL>(ndst[j], pend.id_top[j], pend.top[j], c.c0, C);
          flush_open_p<G, FULL>(ndst[j], pend.id_bot[j], pend.bot[j], c.c0, C);
        }
        g_d += plane_stride;
      }
    }
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int i = 0; i < c.npix; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(i * G + g));
        if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
      }
    }
  }
  tmem_free_cta<kCols>(&s_tmem, warp);
}

// ---------------------------------------------------------------------------
// Slim hand-off kernel with software-pipelined loads (the default since round 1f; also 5 = 16).  sweep_bwd_runf with the
// pipelining of sweep_bwd_runq2 needs ~190 registers (168 + 80 bytes of spills: slower).  Same
// algorithm on a register diet: the hand-off queues are addressed with 32-bit shared-space
// addresses computed from two bases (was six 64-bit pointers), the sample of a pixel is re-read
// from the table where it is used (weights at the blend, weights + offsets at the scatter)
// instead of living in 16 registers across the loads of the next pixel, and the per-pixel
// decisions stay packed in their flag byte.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_a(unsigned a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(unsigned a, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" :: "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void sts_p4(unsigned a, P4 v) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" :: "r"(a), "l"(v.lo), "l"(v.hi) : "memory");
}
__device__ __forceinline__ P4 lds_p4(unsigned a) {
  P4 v;
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_u4(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned lds_u8(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

template <int KMAX>
struct HoLite {
  unsigned slot_out, slot_in;    // shared address of (boundary below / above, neighbour 0, stage 0), this lane's column
  unsigned bar_out, bar_in;      // shared address of the boundary's first {full, empty} barrier pair
  unsigned h_out[KMAX], h_in[KMAX];
};

// one neighbour (J) of one pixel: wq = (w00, w01, w10, w11), oq = (p00, p01, p10, p11)
template <int G, bool FULL, int NSTG, int J, int KMAX>
__device__ __forceinline__ void scatter_hl(float* dst, const P4 (&gw)[G], float4 wq, uint4 oq,
                                           unsigned& id_top, P4 (&top)[G], unsigned& id_bot,
                                           P4 (&bot)[G], unsigned flags, HoLite<KMAX>& ho, int c0, int C) {
  static_assert((NSTG & (NSTG - 1)) == 0, "stage count must be a power of two");
  constexpr unsigned kVec = 512u;                     // 32 lanes x 16 bytes
  constexpr unsigned kSlot = 2u * G * kVec;           // left + right weighted vectors
  constexpr unsigned jslot = (unsigned)J * NSTG * kSlot, jbar = (unsigned)J * NSTG * 16u;
  if (flags & kHoSend) {
    const unsigned h = ho.h_out[J]++;
    const unsigned stg = h & (unsigned)(NSTG - 1), ph = (h / (unsigned)NSTG) & 1u;
    const unsigned bar = ho.bar_out + jbar + stg * 16u;
    mbar_wait_a(bar + 8u, ph ^ 1u);                   // slot empty
    const unsigned a = ho.slot_out + jslot + stg * kSlot;
    const u64 w10 = pk2(wq.z, wq.z), w11 = pk2(wq.w, wq.w);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sts_p4(a + (unsigned)g * kVec, p4scale(gw[g], w10));
      sts_p4(a + (unsigned)(G + g) * kVec, p4scale(gw[g], w11));
    }
    mbar_arrive_a(bar);                               // slot full
  } else {
    side_q<G, FULL>(dst, gw, wq.z, wq.w, oq.z, oq.w, id_bot, bot, c0, C);
  }
  if (flags & kHoRecv) {
    const unsigned h = ho.h_in[J]++;
    const unsigned stg = h & (unsigned)(NSTG - 1), ph = (h / (unsigned)NSTG) & 1u;
    const unsigned bar = ho.bar_in + jbar + stg * 16u;
    mbar_wait_a(bar, ph);
    const unsigned a = ho.slot_in + jslot + stg * kSlot;
    P4 cl[G], cr[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cl[g] = lds_p4(a + (unsigned)g * kVec);
      cr[g] = lds_p4(a + (unsigned)(G + g) * kVec);
    }
    mbar_arrive_a(bar + 8u);
    const u64 w00 = pk2(wq.x, wq.x), w01 = pk2(wq.y, wq.y);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cl[g] = p4fma(gw[g], w00, cl[g]);
      cr[g] = p4fma(gw[g], w01, cr[g]);
    }
    side_c<G, FULL>(dst, cl, cr, (flags & kHoNzLeft) != 0u, (flags & kHoNzRight) != 0u, oq.x, oq.y,
                    id_top, top, c0, C);
  } else {
    side_q<G, FULL>(dst, gw, wq.x, wq.y, oq.x, oq.y, id_top, top, c0, C);
  }
}

template <typename TIn, int G>
__device__ __forceinline__ void blend_taps_w(const RawTaps<TIn, G>& r, float4 wq, P4 (&wv)[G]) {
  WarpSample s;
  s.w00 = wq.x; s.w01 = wq.y; s.w10 = wq.z; s.w11 = wq.w;
  blend_taps<TIn, G>(r, s, wv);
}

// sa: shared address of this pixel's first sample (32 bytes per sample)
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int NSTG, bool V0, bool V1>
__device__ __forceinline__ void pixel_q3(RunPending<KMAX, G>& pend, PixelRaw<TIn, TG, G>& raw, unsigned sa,
                                         unsigned flags0, unsigned flags1, bool has_next,
                                         const WarpSample* smp_next, const TG* __restrict__ gp_next,
                                         const TIn* __restrict__ rp_next, const TIn* const (&nsrc)[KMAX],
                                         float* const (&ndst)[KMAX], uint32_t taddr, u64 inv_n2,
                                         u64 two_inv_n2, int c0, int C, HoLite<KMAX>& ho) {
  constexpr unsigned kS1 = 32u * (KMAX - 1);
  P4 w0[G], w1[G], gw0[G], gw1[G], ref[G], gv[G];
  if (V0) blend_taps_w<TIn, G>(raw.t0, lds_f4(sa), w0);
  if (V1) blend_taps_w<TIn, G>(raw.t1, lds_f4(sa + kS1), w1);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ref[g] = p4from(raw.r[g]);
    gv[g] = p4scale(p4from(raw.g[g]), two_inv_n2);
  }
  if (has_next) issue_pixel_loads<TIn, TG, KMAX, G, FULL>(raw, smp_next, gp_next, rp_next, nsrc, c0, C);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    P4 mu = ref[g];
    if (V0) mu = p4add(mu, w0[g]);
    if (V1) mu = p4add(mu, w1[g]);
    mu = p4scale(mu, inv_n2);
    const uint32_t ta = taddr + 4u * (uint32_t)g;
    tmem_st4(ta, p4fma(gv[g], p4sub(ref[g], mu), tmem_ld4(ta)));
    if (V0) gw0[g] = p4mul(gv[g], p4sub(w0[g], mu));
    if (V1) gw1[g] = p4mul(gv[g], p4sub(w1[g], mu));
  }
  if (V0)
    scatter_hl<G, FULL, NSTG, 0, KMAX>(ndst[0], gw0, lds_f4(sa), lds_u4(sa + 16u), pend.id_top[0], pend.top[0],
                                       pend.id_bot[0], pend.bot[0], flags0, ho, c0, C);
  if (V1)
    scatter_hl<G, FULL, NSTG, KMAX - 1, KMAX>(ndst[KMAX - 1], gw1, lds_f4(sa + kS1), lds_u4(sa + kS1 + 16u),
                                              pend.id_top[KMAX - 1], pend.top[KMAX - 1], pend.id_bot[KMAX - 1],
                                              pend.bot[KMAX - 1], flags1, ho, c0, C);
}

// requires p.k == KMAX
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int MINB, int NSTG>
__global__ void __launch_bounds__(kRunThreads, MINB) sweep_bwd_runs_kernel(const SweepParams p) {
  constexpr int kCols = kRun * G * 4;
  constexpr unsigned kSlot = 2u * G * 512u;
  extern __shared__ __align__(16) unsigned char s_dyn[];          // [kRunRows - 1][KMAX][NSTG] slots
  // two table buffers per warp: while the planes of one are processed the other already holds the
  // next planes, so the load pipeline never drains at a table refill
  __shared__ WarpSample s_tab[kRunRows][2][32];
  __shared__ unsigned char s_flg[kRunRows][2][32];
  __shared__ __align__(8) unsigned long long s_bar[kRunRows - 1][KMAX][NSTG][2];   // {full, empty}
  __shared__ uint32_t s_tmem;
  static_assert(sizeof(WarpSample) == 32, "sample table stride");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  if (threadIdx.x < (kRunRows - 1) * KMAX * NSTG * 2) {
    mbar_init(&s_bar[0][0][0][0] + threadIdx.x, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);     // contains the CTA barriers
  if (c.y < p.H) {
    const int C = p.C, HW = p.H * p.W;
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const TIn* ref_row = feat + ref_off;
    const size_t plane_stride = (size_t)HW * C;
    const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const bool one_chunk = p.slices == 1;
    const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                         sizeof(TG)) & ~15u;
    const TG* pf_base = g_d - 4 * lane;
    const bool pf_ok = pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                       ((plane_stride * sizeof(TG)) & 15) == 0;
    auto prefetch_plane = [&](int d) {
      if (!pf_ok || d >= p.D) return;
      const TG* q = pf_base + (size_t)d * plane_stride;
      if (one_chunk) {
        if (lane == 0) prefetch_l2(q, pf_bytes);
      } else if (lane < c.npix) {
        prefetch_l2(q + (size_t)lane * C, pf_bytes);
      }
    };
#pragma unroll
    for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int n = __ldg(p.nbr + (size_t)c.v * KMAX + j);
      nsrc[j] = feat + (size_t)n * HW * C + c.c0;
      ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(KMAX + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
    constexpr int spp = kRun * KMAX;
    constexpr int ppf = 32 / spp > 0 ? 32 / spp : 1;
    const bool has_up = warp > 0;
    const bool has_dn = warp + 1 < kRunRows && c.y + 1 < p.H;
    const int bo = min(warp, kRunRows - 2), bi = max(warp - 1, 0);     // boundary below / above this row
    HoLite<KMAX> ho;
    ho.slot_out = smem_u32(s_dyn) + (unsigned)(bo * KMAX * NSTG) * kSlot + (unsigned)lane * 16u;
    ho.slot_in = smem_u32(s_dyn) + (unsigned)(bi * KMAX * NSTG) * kSlot + (unsigned)lane * 16u;
    ho.bar_out = smem_u32(&s_bar[bo][0][0][0]);
    ho.bar_in = smem_u32(&s_bar[bi][0][0][0]);
#pragma unroll
    for (int j = 0; j < KMAX; ++j) ho.h_out[j] = ho.h_in[j] = 0u;
    const unsigned tab_a = smem_u32(s_tab[warp][0]), flg_a = smem_u32(s_flg[warp][0]);

#pragma unroll
    for (int q = 0; q < kRun * G; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();

    PixelRaw<TIn, TG, G> raw;
    fill_run_samples_ho(s_tab[warp][0], s_flg[warp][0], p, c, 0, ppf, lane, has_up, has_dn);
    if (ppf < p.D) fill_run_samples_ho(s_tab[warp][1], s_flg[warp][1], p, c, ppf, ppf, lane, has_up, has_dn);
    __syncwarp();
    // pipeline prologue: first pixel of the first plane
    issue_pixel_loads<TIn, TG, KMAX, G, FULL>(raw, s_tab[warp][0], g_d, ref_row, nsrc, c.c0, C);
    int buf = 0;
    for (int d0 = 0; d0 < p.D; d0 += ppf, buf ^= 1) {
      const int dend = min(p.D, d0 + ppf);
      const bool more_fills = d0 + ppf < p.D;
      for (int d = d0; d < dend; ++d) {
        prefetch_plane(d + kPrefetchPlanes);
        tmem_wait_st();
        RunPending<KMAX, G> pend;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          pend.id_top[j] = pend.id_bot[j] = kNoTap;
#pragma unroll
          for (int g = 0; g < G; ++g) pend.top[j][g] = pend.bot[j][g] = p4zero();
        }
        const int toff = buf * 32 + (d - d0) * spp;
        const WarpSample* tab = s_tab[warp][0] + toff;
        const bool last_plane = d + 1 >= dend;
        const bool more_planes = !last_plane || more_fills;
        // first sample of the next plane: next rows of this buffer, or the other buffer
        const WarpSample* tab_next = last_plane ? s_tab[warp][buf ^ 1] : tab + spp;
#pragma unroll 1
        for (int i = 0; i < c.npix; ++i) {
          const unsigned si = (unsigned)(toff + i * KMAX);
          const unsigned sa = tab_a + si * 32u;
          const bool v0 = lds_u32(sa + 16u) != kNoSample;
          const bool v1 = KMAX == 2 && lds_u32(sa + 32u * (KMAX - 1) + 16u) != kNoSample;
          const unsigned f0 = lds_u8(flg_a + si);
          const unsigned f1 = KMAX == 2 ? lds_u8(flg_a + si + (KMAX - 1)) : 0u;
          const bool in_run = i + 1 < c.npix;
          const bool has_next = in_run || more_planes;
          const WarpSample* smp_next = in_run ? tab + (i + 1) * KMAX : tab_next;
          const TG* gp_next = in_run ? g_d + (i + 1) * C : g_d + plane_stride;
          const TIn* rp_next = in_run ? ref_row + (i + 1) * C : ref_row;
          const uint32_t ta = tbase + 4u * (uint32_t)(i * G);
          if (v0 && v1)
            pixel_q3<TIn, TG, KMAX, G, FULL, NSTG, true, true>(pend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
          else if (v0)
            pixel_q3<TIn, TG, KMAX, G, FULL, NSTG, true, false>(pend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
          else if (v1)
            pixel_q3<TIn, TG, KMAX, G, FULL, NSTG, false, true>(pend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
          else
            pixel_q3<TIn, TG, KMAX, G, FULL, NSTG, false, false>(pend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          flush_open_p<G, FULL>(ndst[j], pend.id_top[j], pend.top[j], c.c0, C);
          flush_open_p<G, FULL>(ndst[j], pend.id_bot[j], pend.bot[j], c.c0, C);
        }
        g_d += plane_stride;
      }
      // this buffer's planes are done (the loads already in flight read the other buffer): refill it
      __syncwarp();
      if (d0 + 2 * ppf < p.D)
        fill_run_samples_ho(s_tab[warp][buf], s_flg[warp][buf], p, c, d0 + 2 * ppf, ppf, lane, has_up, has_dn);
      __syncwarp();
    }
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int i = 0; i < c.npix; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(i * G + g));
        if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
      }
    }
  }
  tmem_free_cta<kCols>(&s_tmem, warp);
}

#ifdef MVSD_EXP_TMEM_PENDING
// ---------------------------------------------------------------------------
// EXPERIMENT, NOT BUILT BY DEFAULT AND NOT YET RUN ON A GPU: the slim
// hand-off kernel with the pending taps in tensor memory instead of 32 registers, compiled for four
// CTAs per SM (128 registers).  Written at the end of round 1 to answer the register question with
// ptxas; correctness and speed are the first measurement of round 2
// (build with MVSD_EXTRA_NVCC_FLAGS=-DMVSD_EXP_TMEM_PENDING, select with tuning key 5 = 17).
// TMEM columns per CTA: [0, 64) reference gradients, [64, 96) pending taps
// ((neighbour * 2 + side) * G + group) * 4; allocation 128 columns, 4 CTAs = the whole TMEM.
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void tm_ldv(uint32_t ta, P4 (&v)[G]) {
  float f[G][4];
#pragma unroll
  for (int g = 0; g < G; ++g)
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(f[g][0]), "=f"(f[g][1]), "=f"(f[g][2]), "=f"(f[g][3]) : "r"(ta + 4u * (uint32_t)g) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int g = 0; g < G; ++g) v[g] = P4{pk2(f[g][0], f[g][1]), pk2(f[g][2], f[g][3])};
}
template <int G>
__device__ __forceinline__ void tm_stv(uint32_t ta, const P4 (&v)[G]) {
#pragma unroll
  for (int g = 0; g < G; ++g) tmem_st4(ta + 4u * (uint32_t)g, v[g]);
}

// side_q with the pending accumulator of this (neighbour, side) at TMEM address ta
template <int G, bool FULL>
__device__ __forceinline__ void side_qt(float* dst, const P4 (&gw)[G], float w_left, float w_right,
                                        unsigned p_left, unsigned p_right, unsigned& open_id, uint32_t ta,
                                        int c0, int C) {
  const u64 wl = pk2(w_left, w_left), wr = pk2(w_right, w_right);
  P4 a[G], o[G];
  tmem_wait_st();                          // the previous pixel's store to these columns
  if (open_id == p_left) {
    tm_ldv<G>(ta, o);
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = p4fma(gw[g], wl, o[g]);
    red_group_p<G, FULL>(dst, p_left, a, c0, C);
#pragma unroll
    for (int g = 0; g < G; ++g) o[g] = p4scale(gw[g], wr);
    tm_stv<G>(ta, o);
    open_id = w_right != 0.f ? p_right : kNoTap;
  } else if (open_id == p_right && w_right != 0.f) {
    tm_ldv<G>(ta, o);
#pragma unroll
    for (int g = 0; g < G; ++g) o[g] = p4fma(gw[g], wr, o[g]);
    tm_stv<G>(ta, o);
    if (w_left != 0.f) {
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = p4scale(gw[g], wl);
      red_group_p<G, FULL>(dst, p_left, a, c0, C);
    }
  } else {
    if (open_id != kNoTap) {
      tm_ldv<G>(ta, o);
      red_group_p<G, FULL>(dst, open_id, o, c0, C);
    }
    if (w_left != 0.f) {
#pragma unroll
      for (int g = 0; g < G; ++g) a[g] = p4scale(gw[g], wl);
      red_group_p<G, FULL>(dst, p_left, a, c0, C);
    }
#pragma unroll
    for (int g = 0; g < G; ++g) o[g] = p4scale(gw[g], wr);
    tm_stv<G>(ta, o);
    open_id = w_right != 0.f ? p_right : kNoTap;
  }
}

// side_c (pre-weighted left / right contributions) with the pending accumulator in TMEM
template <int G, bool FULL>
__device__ __forceinline__ void side_ct(float* dst, const P4 (&cl)[G], const P4 (&cr)[G], bool nz_left,
                                        bool nz_right, unsigned p_left, unsigned p_right, unsigned& open_id,
                                        uint32_t ta, int c0, int C) {
  P4 a[G], o[G];
  tmem_wait_st();
  if (open_id == p_left) {
    tm_ldv<G>(ta, o);
#pragma unroll
    for (int g = 0; g < G; ++g) a[g] = p4add(cl[g], o[g]);
    red_group_p<G, FULL>(dst, p_left, a, c0, C);
    tm_stv<G>(ta, cr);
    open_id = nz_right ? p_right : kNoTap;
  } else if (open_id == p_right && nz_right) {
    tm_ldv<G>(ta, o);
#pragma unroll
    for (int g = 0; g < G; ++g) o[g] = p4add(cr[g], o[g]);
    tm_stv<G>(ta, o);
    if (nz_left) red_group_p<G, FULL>(dst, p_left, cl, c0, C);
  } else {
    if (open_id != kNoTap) {
      tm_ldv<G>(ta, o);
      red_group_p<G, FULL>(dst, open_id, o, c0, C);
    }
    if (nz_left) red_group_p<G, FULL>(dst, p_left, cl, c0, C);
    tm_stv<G>(ta, cr);
    open_id = nz_right ? p_right : kNoTap;
  }
}

template <int G, bool FULL>
__device__ __forceinline__ void flush_open_t(float* dst, unsigned& id, uint32_t ta, int c0, int C) {
  if (id != kNoTap) {
    P4 o[G];
    tmem_wait_st();
    tm_ldv<G>(ta, o);
    red_group_p<G, FULL>(dst, id, o, c0, C);
  }
  id = kNoTap;
}

template <int KMAX>
struct PendingIds {
  unsigned top[KMAX], bot[KMAX];
};

template <int G, bool FULL, int NSTG, int J, int KMAX>
__device__ __forceinline__ void scatter_ht(float* dst, const P4 (&gw)[G], float4 wq, uint4 oq,
                                           unsigned& id_top, unsigned& id_bot, uint32_t ta_top, uint32_t ta_bot,
                                           unsigned flags, HoLite<KMAX>& ho, int c0, int C) {
  constexpr unsigned kVec = 512u;
  constexpr unsigned kSlot = 2u * G * kVec;
  constexpr unsigned jslot = (unsigned)J * NSTG * kSlot, jbar = (unsigned)J * NSTG * 16u;
  if (flags & kHoSend) {
    const unsigned h = ho.h_out[J]++;
    const unsigned stg = h & (unsigned)(NSTG - 1), ph = (h / (unsigned)NSTG) & 1u;
    const unsigned bar = ho.bar_out + jbar + stg * 16u;
    mbar_wait_a(bar + 8u, ph ^ 1u);
    const unsigned a = ho.slot_out + jslot + stg * kSlot;
    const u64 w10 = pk2(wq.z, wq.z), w11 = pk2(wq.w, wq.w);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sts_p4(a + (unsigned)g * kVec, p4scale(gw[g], w10));
      sts_p4(a + (unsigned)(G + g) * kVec, p4scale(gw[g], w11));
    }
    mbar_arrive_a(bar);
  } else {
    side_qt<G, FULL>(dst, gw, wq.z, wq.w, oq.z, oq.w, id_bot, ta_bot, c0, C);
  }
  if (flags & kHoRecv) {
    const unsigned h = ho.h_in[J]++;
    const unsigned stg = h & (unsigned)(NSTG - 1), ph = (h / (unsigned)NSTG) & 1u;
    const unsigned bar = ho.bar_in + jbar + stg * 16u;
    mbar_wait_a(bar, ph);
    const unsigned a = ho.slot_in + jslot + stg * kSlot;
    P4 cl[G], cr[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cl[g] = lds_p4(a + (unsigned)g * kVec);
      cr[g] = lds_p4(a + (unsigned)(G + g) * kVec);
    }
    mbar_arrive_a(bar + 8u);
    const u64 w00 = pk2(wq.x, wq.x), w01 = pk2(wq.y, wq.y);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      cl[g] = p4fma(gw[g], w00, cl[g]);
      cr[g] = p4fma(gw[g], w01, cr[g]);
    }
    side_ct<G, FULL>(dst, cl, cr, (flags & kHoNzLeft) != 0u, (flags & kHoNzRight) != 0u, oq.x, oq.y, id_top,
                     ta_top, c0, C);
  } else {
    side_qt<G, FULL>(dst, gw, wq.x, wq.y, oq.x, oq.y, id_top, ta_top, c0, C);
  }
}

template <typename TIn, typename TG, int KMAX, int G, bool FULL, int NSTG, bool V0, bool V1>
__device__ __forceinline__ void pixel_q4(PendingIds<KMAX>& ids, uint32_t tpend, PixelRaw<TIn, TG, G>& raw,
                                         unsigned sa, unsigned flags0, unsigned flags1, bool has_next,
                                         const WarpSample* smp_next, const TG* __restrict__ gp_next,
                                         const TIn* __restrict__ rp_next, const TIn* const (&nsrc)[KMAX],
                                         float* const (&ndst)[KMAX], uint32_t taddr, u64 inv_n2,
                                         u64 two_inv_n2, int c0, int C, HoLite<KMAX>& ho) {
  constexpr unsigned kS1 = 32u * (KMAX - 1);
  constexpr uint32_t kSide = 4u * G;                  // TMEM columns of one pending accumulator
  P4 w0[G], w1[G], gw0[G], gw1[G], ref[G], gv[G];
  if (V0) blend_taps_w<TIn, G>(raw.t0, lds_f4(sa), w0);
  if (V1) blend_taps_w<TIn, G>(raw.t1, lds_f4(sa + kS1), w1);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ref[g] = p4from(raw.r[g]);
    gv[g] = p4scale(p4from(raw.g[g]), two_inv_n2);
  }
  if (has_next) issue_pixel_loads<TIn, TG, KMAX, G, FULL>(raw, smp_next, gp_next, rp_next, nsrc, c0, C);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    P4 mu = ref[g];
    if (V0) mu = p4add(mu, w0[g]);
    if (V1) mu = p4add(mu, w1[g]);
    mu = p4scale(mu, inv_n2);
    const uint32_t ta = taddr + 4u * (uint32_t)g;
    tmem_st4(ta, p4fma(gv[g], p4sub(ref[g], mu), tmem_ld4(ta)));
    if (V0) gw0[g] = p4mul(gv[g], p4sub(w0[g], mu));
    if (V1) gw1[g] = p4mul(gv[g], p4sub(w1[g], mu));
  }
  if (V0)
    scatter_ht<G, FULL, NSTG, 0, KMAX>(ndst[0], gw0, lds_f4(sa), lds_u4(sa + 16u), ids.top[0], ids.bot[0],
                                       tpend, tpend + kSide, flags0, ho, c0, C);
  if (V1)
    scatter_ht<G, FULL, NSTG, KMAX - 1, KMAX>(ndst[KMAX - 1], gw1, lds_f4(sa + kS1), lds_u4(sa + kS1 + 16u),
                                              ids.top[KMAX - 1], ids.bot[KMAX - 1],
                                              tpend + 2u * kSide * (KMAX - 1), tpend + 2u * kSide * (KMAX - 1) + kSide,
                                              flags1, ho, c0, C);
}

// requires p.k == KMAX
template <typename TIn, typename TG, int KMAX, int G, bool FULL, int NSTG>
__global__ void __launch_bounds__(kRunThreads, 4) sweep_bwd_runt_kernel(const SweepParams p) {
  constexpr int kRefCols = kRun * G * 4;                       // 64
  constexpr int kCols = 128;                                   // 64 + 2 * KMAX * G * 4 <= 96 -> 128
  static_assert(kRefCols + 2 * KMAX * G * 4 <= kCols, "TMEM budget");
  constexpr unsigned kSlot = 2u * G * 512u;
  extern __shared__ __align__(16) unsigned char s_dyn[];
  __shared__ WarpSample s_tab[kRunRows][2][32];
  __shared__ unsigned char s_flg[kRunRows][2][32];
  __shared__ __align__(8) unsigned long long s_bar[kRunRows - 1][KMAX][NSTG][2];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RunCoord c = run_coord<G>(p, warp, lane);
  if (threadIdx.x < (kRunRows - 1) * KMAX * NSTG * 2) {
    mbar_init(&s_bar[0][0][0][0] + threadIdx.x, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tbase = tmem_alloc_cta<kCols>(&s_tmem, warp);
  const uint32_t tpend = tbase + (uint32_t)kRefCols;
  if (c.y < p.H) {
    const int C = p.C, HW = p.H * p.W;
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t ref_off = ((size_t)(c.v + p.ref_begin) * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const TIn* ref_row = feat + ref_off;
    const size_t plane_stride = (size_t)HW * C;
    const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)c.v * p.D * HW + (size_t)c.y * p.W + c.x0) * C + c.c0;
    const bool one_chunk = p.slices == 1;
    const unsigned pf_bytes = (unsigned)((one_chunk ? (size_t)c.npix * C : (size_t)min(128 * G, C - (c.c0 - 4 * lane))) *
                                         sizeof(TG)) & ~15u;
    const TG* pf_base = g_d - 4 * lane;
    const bool pf_ok = pf_bytes >= 16 && (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                       ((plane_stride * sizeof(TG)) & 15) == 0;
    auto prefetch_plane = [&](int d) {
      if (!pf_ok || d >= p.D) return;
      const TG* q = pf_base + (size_t)d * plane_stride;
      if (one_chunk) {
        if (lane == 0) prefetch_l2(q, pf_bytes);
      } else if (lane < c.npix) {
        prefetch_l2(q + (size_t)lane * C, pf_bytes);
      }
    };
#pragma unroll
    for (int d = 0; d < kPrefetchPlanes; ++d) prefetch_plane(d);

    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int n = __ldg(p.nbr + (size_t)c.v * KMAX + j);
      nsrc[j] = feat + (size_t)n * HW * C + c.c0;
      ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(KMAX + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);
    constexpr int spp = kRun * KMAX;
    constexpr int ppf = 32 / spp > 0 ? 32 / spp : 1;
    const bool has_up = warp > 0;
    const bool has_dn = warp + 1 < kRunRows && c.y + 1 < p.H;
    const int bo = min(warp, kRunRows - 2), bi = max(warp - 1, 0);
    HoLite<KMAX> ho;
    ho.slot_out = smem_u32(s_dyn) + (unsigned)(bo * KMAX * NSTG) * kSlot + (unsigned)lane * 16u;
    ho.slot_in = smem_u32(s_dyn) + (unsigned)(bi * KMAX * NSTG) * kSlot + (unsigned)lane * 16u;
    ho.bar_out = smem_u32(&s_bar[bo][0][0][0]);
    ho.bar_in = smem_u32(&s_bar[bi][0][0][0]);
#pragma unroll
    for (int j = 0; j < KMAX; ++j) ho.h_out[j] = ho.h_in[j] = 0u;
    const unsigned tab_a = smem_u32(s_tab[warp][0]), flg_a = smem_u32(s_flg[warp][0]);

#pragma unroll
    for (int q = 0; q < kRun * G; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();

    PixelRaw<TIn, TG, G> raw;
    fill_run_samples_ho(s_tab[warp][0], s_flg[warp][0], p, c, 0, ppf, lane, has_up, has_dn);
    if (ppf < p.D) fill_run_samples_ho(s_tab[warp][1], s_flg[warp][1], p, c, ppf, ppf, lane, has_up, has_dn);
    __syncwarp();
    issue_pixel_loads<TIn, TG, KMAX, G, FULL>(raw, s_tab[warp][0], g_d, ref_row, nsrc, c.c0, C);
    int buf = 0;
    for (int d0 = 0; d0 < p.D; d0 += ppf, buf ^= 1) {
      const int dend = min(p.D, d0 + ppf);
      const bool more_fills = d0 + ppf < p.D;
      for (int d = d0; d < dend; ++d) {
        prefetch_plane(d + kPrefetchPlanes);
        tmem_wait_st();
        PendingIds<KMAX> ids;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) ids.top[j] = ids.bot[j] = kNoTap;
        const int toff = buf * 32 + (d - d0) * spp;
        const WarpSample* tab = s_tab[warp][0] + toff;
        const bool last_plane = d + 1 >= dend;
        const bool more_planes = !last_plane || more_fills;
        const WarpSample* tab_next = last_plane ? s_tab[warp][buf ^ 1] : tab + spp;
#pragma unroll 1
        for (int i = 0; i < c.npix; ++i) {
          const unsigned si = (unsigned)(toff + i * KMAX);
          const unsigned sa = tab_a + si * 32u;
          const bool v0 = lds_u32(sa + 16u) != kNoSample;
          const bool v1 = KMAX == 2 && lds_u32(sa + 32u * (KMAX - 1) + 16u) != kNoSample;
          const unsigned f0 = lds_u8(flg_a + si);
          const unsigned f1 = KMAX == 2 ? lds_u8(flg_a + si + (KMAX - 1)) : 0u;
          const bool in_run = i + 1 < c.npix;
          const bool has_next = in_run || more_planes;
          const WarpSample* smp_next = in_run ? tab + (i + 1) * KMAX : tab_next;
          const TG* gp_next = in_run ? g_d + (i + 1) * C : g_d + plane_stride;
          const TIn* rp_next = in_run ? ref_row + (i + 1) * C : ref_row;
          const uint32_t ta = tbase + 4u * (uint32_t)(i * G);
          if (v0 && v1)
            pixel_q4<TIn, TG, KMAX, G, FULL, NSTG, true, true>(ids, tpend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
          else if (v0)
            pixel_q4<TIn, TG, KMAX, G, FULL, NSTG, true, false>(ids, tpend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
          else if (v1)
            pixel_q4<TIn, TG, KMAX, G, FULL, NSTG, false, true>(ids, tpend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
          else
            pixel_q4<TIn, TG, KMAX, G, FULL, NSTG, false, false>(ids, tpend, raw, sa, f0, f1, has_next, smp_next, gp_next, rp_next, nsrc, ndst, ta, inv_n2, two_inv_n2, c.c0, C, ho);
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          flush_open_t<G, FULL>(ndst[j], ids.top[j], tpend + 8u * G * (uint32_t)j, c.c0, C);
          flush_open_t<G, FULL>(ndst[j], ids.bot[j], tpend + 8u * G * (uint32_t)j + 4u * G, c.c0, C);
        }
        g_d += plane_stride;
      }
      __syncwarp();
      if (d0 + 2 * ppf < p.D)
        fill_run_samples_ho(s_tab[warp][buf], s_flg[warp][buf], p, c, d0 + 2 * ppf, ppf, lane, has_up, has_dn);
      __syncwarp();
    }
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int i = 0; i < c.npix; ++i) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(i * G + g));
        if (group_on<FULL>(c.c0, g, C)) red_add_p4(dst + i * C + 128 * g, acc);
      }
    }
  }
  tmem_free_cta<kCols>(&s_tmem, warp);
}
#endif  // MVSD_EXP_TMEM_PENDING

// k in {1,2} only (k*kRun <= 32 samples per plane); other k use the pixel kernel.
template <typename TIn, typename TG>
static int launch_bwd_run_t(SweepParams& p, cudaStream_t st) {
  const int G = sweep_groups(p.C);
  p.tiles_x = (p.W + kRun - 1) / kRun;
  p.tiles_y = (p.H + kRunRows - 1) / kRunRows;
  p.slices = (p.C + 128 * G - 1) / (128 * G);
  const long long blocks = (long long)p.V * p.slices * p.tiles_y * p.tiles_x;
  if (blocks > 2147483647LL) return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: grid too large");
  dim3 grid((unsigned)blocks);
  const bool full = p.C % (128 * G) == 0;
  const bool tm = tuning(4) == 0;          // tuning key 4: 0 = TMEM accumulators (default), 2 = shared memory
  const bool minb3 = tuning(4) != 1;       // tuning key 4: 1 = cap at 128 registers (4 CTAs/SM) instead of 168 (3)
  const bool packed = tuning(5) != 2;      // tuning key 5: 2 = scalar-math run kernel, 3 = packed
  // default: slim row hand-off with software-pipelined loads (sweep_bwd_runs) for bf16 features; with
  // fp32 features the raw loads of a pixel are 80 registers instead of 44 and the pipelined kernels
  // spill (ptxas: 160 B of stack for sweep_bwd_runs<float>), so their default stays the un-pipelined
  // lean kernel (sweep_bwd_runq, 168 registers, no spills).  5=16 forces sweep_bwd_runs, 5=14 the
  // pipelined lean kernel without the hand-off (sweep_bwd_runq2), 5=7 sweep_bwd_runq, 5=3 the first
  // packed kernel (sweep_bwd_runp).
  const bool lean = tuning(5) == 0 && tuning(4) == 0;
  const bool slim_ok = sizeof(TIn) == 2;
#ifdef MVSD_EXP_NORED
  { const int flag = tuning(6); cudaMemcpyToSymbolAsync(c_exp_nored, &flag, sizeof(int), 0, cudaMemcpyHostToDevice, st); }
#endif
#if MVSD_KRUN == 8
#define MVSD_RUNH(KM, GG, FU, NS, NVV)                                                      \
  do {                                                                                    \
    auto kern = sweep_bwd_runh_kernel<TIn, TG, KM, GG, FU, 3, NS, NVV>;                    \
    constexpr size_t dyn = runh_slot_bytes<KM, GG, NS, NVV>();                            \
    static bool attr_set = false;                                                         \
    if (!attr_set) {                                                                      \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);  \
      attr_set = true;                                                                    \
    }                                                                                     \
    kern<<<grid, kRunThreads, dyn, st>>>(p);                                              \
  } while (0)
#ifdef MVSD_EXP_TMEM_PENDING
#define MVSD_RUNT(KM, GG, FU, NS)                                                           \
  do {                                                                                    \
    auto kern = sweep_bwd_runt_kernel<TIn, TG, KM, GG, FU, NS>;                            \
    constexpr size_t dyn = runh_slot_bytes<KM, GG, NS, 2>();                              \
    static bool attr_set = false;                                                         \
    if (!attr_set) {                                                                      \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);  \
      attr_set = true;                                                                    \
    }                                                                                     \
    kern<<<grid, kRunThreads, dyn, st>>>(p);                                              \
  } while (0)
#else
#define MVSD_RUNT(KM, GG, FU, NS) MVSD_RUNS(KM, GG, FU, NS)      /* experiment not built: the default */
#endif
#define MVSD_RUNS(KM, GG, FU, NS)                                                           \
  do {                                                                                    \
    auto kern = sweep_bwd_runs_kernel<TIn, TG, KM, GG, FU, 3, NS>;                         \
    constexpr size_t dyn = runh_slot_bytes<KM, GG, NS, 2>();                              \
    static bool attr_set = false;                                                         \
    if (!attr_set) {                                                                      \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);  \
      attr_set = true;                                                                    \
    }                                                                                     \
    kern<<<grid, kRunThreads, dyn, st>>>(p);                                              \
  } while (0)
#define MVSD_RUNF(KM, GG, FU, NS)                                                           \
  do {                                                                                    \
    auto kern = sweep_bwd_runf_kernel<TIn, TG, KM, GG, FU, 3, NS>;                         \
    constexpr size_t dyn = runh_slot_bytes<KM, GG, NS, 2>();                              \
    static bool attr_set = false;                                                         \
    if (!attr_set) {                                                                      \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);  \
      attr_set = true;                                                                    \
    }                                                                                     \
    kern<<<grid, kRunThreads, dyn, st>>>(p);                                              \
  } while (0)
#define MVSD_RUN(KM, GG, FU)                                                              \
  do {                                                                                    \
    if (tuning(5) == 8) MVSD_RUNH(KM, GG, FU, 2, 2);                                         \
    else if (tuning(5) == 11) MVSD_RUNF(KM, GG, FU, 2);                                      \
    else if (tuning(5) == 17) MVSD_RUNT(KM, GG, FU, 2);                                      \
    else if ((lean && slim_ok) || tuning(5) == 16) MVSD_RUNS(KM, GG, FU, 2);                 \
    else if (tuning(5) == 14) sweep_bwd_runq2_kernel<TIn, TG, KM, GG, FU, MVSD_RUNQ_MINB><<<grid, kRunThreads, 0, st>>>(p); \
    else if (lean || tuning(5) == 7) sweep_bwd_runq_kernel<TIn, TG, KM, GG, FU, MVSD_RUNQ_MINB><<<grid, kRunThreads, 0, st>>>(p); \
    else if (packed && tm) sweep_bwd_runp_kernel<TIn, TG, KM, GG, FU, 3, true><<<grid, kRunThreads, 0, st>>>(p); \
    else if (packed && minb3) sweep_bwd_runp_kernel<TIn, TG, KM, GG, FU, 3, false><<<grid, kRunThreads, 0, st>>>(p); \
    else if (packed) sweep_bwd_runp_kernel<TIn, TG, KM, GG, FU, 4, false><<<grid, kRunThreads, 0, st>>>(p); \
    else sweep_bwd_run_kernel<TIn, TG, KM, GG, FU><<<grid, kRunThreads, 0, st>>>(p);        \
  } while (0)
#else   // experiment builds with longer runs: TMEM variant only (the smem accumulators exceed 48 KB)
#define MVSD_RUN(KM, GG, FU) \
  sweep_bwd_runp_kernel<TIn, TG, KM, GG, FU, 3, true><<<grid, kRunThreads, 0, st>>>(p)
#endif
  if (p.k == 1) {
    if (G == 2) { if (full) MVSD_RUN(1, 2, true); else MVSD_RUN(1, 2, false); }
    else { if (full) MVSD_RUN(1, 1, true); else MVSD_RUN(1, 1, false); }
  } else {
    if (G == 2) { if (full) MVSD_RUN(2, 2, true); else MVSD_RUN(2, 2, false); }
    else { if (full) MVSD_RUN(2, 1, true); else MVSD_RUN(2, 1, false); }
  }
#undef MVSD_RUN
#undef MVSD_RUNH
#undef MVSD_RUNF
#undef MVSD_RUNS
#undef MVSD_RUNT
  count_launch();
  return check_launch("plane_sweep_bwd(run)");
}

int launch_bwd_run(SweepParams& p, int feat_dtype, int g_dtype, cudaStream_t st) {
#ifndef MVSD_DEV_FAST      // register-count iterations: compile the headline dtype combination only
  if (feat_dtype == MVSD_F32 && g_dtype == MVSD_F32) return launch_bwd_run_t<float, float>(p, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_BF16)
    return launch_bwd_run_t<__nv_bfloat16, __nv_bfloat16>(p, st);
#endif
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_F32)
    return launch_bwd_run_t<__nv_bfloat16, float>(p, st);
  return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: dtype combination not built");
}

}  // namespace mvsd
