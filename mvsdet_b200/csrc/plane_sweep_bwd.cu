// Plane-sweep variance volume, backward (and the homography warp's backward).
// See plane_sweep.cuh for the design.  What autograd computes through
// mvsdet.py:439-467, with n = k+1, mu = S1/n:
//   dL/dref       += sum_d G (2/n) (ref - mu_d)
//   dL/dwarped_j   = G (2/n) (warped_j - mu)   -> bilinear scatter into neighbour j
// The scatter uses fp32 vector REDs (the L2 atomic unit is the binding
// resource: 32 B / clk / slice); along a run the contributions to a tap column
// shared by two consecutive pixels are summed in registers first, so a source
// pixel receives one RED per run instead of two.
#include "plane_sweep.cuh"

namespace mvsd {

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

// One side (top or bottom row) of the scatter of one sample: the left tap
// merges with the pending right tap of the previous pixel when it is the same
// source pixel and leaves as one RED; the right tap stays pending.
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

#ifndef MVSD_BWD_MINB
#define MVSD_BWD_MINB 1
#endif
template <typename TIn, typename TG, int KMAX, int G, bool FULL, bool WARP_ONLY>
__global__ void __launch_bounds__(kSweepThreads, MVSD_BWD_MINB) sweep_bwd_kernel(const SweepParams p) {
  __shared__ WarpSample s_tab[kRows][32];
  __shared__ float4 s_gref[WARP_ONLY ? 1 : kRows][WARP_ONLY ? 1 : kRun][G][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const SweepCoord c = sweep_coord<G>(p, warp, lane);
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
    if (!WARP_ONLY && j < k) n = __ldg(p.nbr + (size_t)c.v * k + j);
    nsrc[j] = feat + (size_t)n * HW * C + c.c0;
    ndst[j] = p.g_feat + (size_t)n * HW * C + c.c0;
  }
  const float inv_n = 1.0f / (float)(k + 1);
  const float two_inv_n = 2.0f * inv_n;
  const int spp = kRun * k;
  const int ppf = k > 0 ? max(1, 32 / spp) : p.D;

  if (!WARP_ONLY) {
#pragma unroll
    for (int i = 0; i < kRun; ++i)
#pragma unroll
      for (int g = 0; g < G; ++g) s_gref[warp][i][g][lane] = f4zero();
  }

  for (int d0 = 0; d0 < p.D; d0 += ppf) {
    if (k > 0) {
      __syncwarp();
      fill_samples(s_tab[warp], p, c, d0, ppf, lane);
      __syncwarp();
    }
    const int dend = min(p.D, d0 + ppf);
    for (int d = d0; d < dend; ++d) {
      float4 open_top[KMAX][G], open_bot[KMAX][G];   // pending right-column contributions
      unsigned o_top[KMAX], o_bot[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        o_top[j] = o_bot[j] = kNoTap;
#pragma unroll
        for (int g = 0; g < G; ++g) open_top[j][g] = open_bot[j][g] = f4zero();
      }
      const WarpSample* tab = s_tab[warp] + (d - d0) * spp;
      const TG* g_d = g_row + (size_t)d * HW * C;
#pragma unroll
      for (int i = 0; i < kRun; ++i) {
        if (i >= c.npix) break;
        float4 gv[G], ref[G], mu[G];
        float4 wv[KMAX][G];
        WarpSample smp[KMAX];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const bool on = group_on<FULL>(c.c0, g, C);
          gv[g] = on ? Io<TG>::ld_stream(g_d + i * C + 128 * g) : f4zero();
          ref[g] = (!WARP_ONLY && on) ? Io<TIn>::ld(ref_row + i * C + 128 * g) : f4zero();
          mu[g] = ref[g];
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
#pragma unroll
          for (int g = 0; g < G; ++g) wv[j][g] = f4zero();
          if (j >= k) continue;
          smp[j] = tab[i * k + j];
          if (WARP_ONLY || smp[j].p00 == kNoSample) continue;
          float4 col[2][2][G];
          unsigned t0 = kNoTap, t1 = kNoTap;
          gather_taps<TIn, G, FULL, false>(nsrc[j], smp[j], c.c0, C, col, t0, t1, i, wv[j]);
#pragma unroll
          for (int g = 0; g < G; ++g) mu[g] = f4add(mu[g], wv[j][g]);
        }
        if (!WARP_ONLY) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            mu[g] = f4scale(mu[g], inv_n);
            gv[g] = f4scale(gv[g], two_inv_n);
            s_gref[warp][i][g][lane] = f4fma(gv[g], f4sub(ref[g], mu[g]), s_gref[warp][i][g][lane]);
          }
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
          for (int g = 0; g < G; ++g)
            gw[g] = WARP_ONLY ? gv[g] : f4mul(gv[g], f4sub(wv[j][g], mu[g]));
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
  if (!WARP_ONLY) {
    float* dst = p.g_feat + ref_off;
#pragma unroll
    for (int i = 0; i < kRun; ++i) {
      if (i >= c.npix) break;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (group_on<FULL>(c.c0, g, C)) red_add_f32x4(dst + i * C + 128 * g, s_gref[warp][i][g][lane]);
    }
  }
}

template <typename TIn, typename TG, bool WARP_ONLY>
static int launch_bwd_k(SweepParams& p, cudaStream_t st) {
  dim3 grid;
  const int G = sweep_groups(p.C);
  if (!sweep_grid(p, G, grid)) return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: grid too large");
  const bool full = p.C % (128 * G) == 0;
  const int kmax = WARP_ONLY ? 1 : (p.k <= 1 ? 1 : (p.k == 2 ? 2 : 4));
#define MVSD_BWD(KM, GG, FU) \
  sweep_bwd_kernel<TIn, TG, KM, GG, FU, WARP_ONLY><<<grid, kSweepThreads, 0, st>>>(p)
#define MVSD_BWD_G(KM)                                                            \
  do {                                                                            \
    if (G == 2) { if (full) MVSD_BWD(KM, 2, true); else MVSD_BWD(KM, 2, false); } \
    else { if (full) MVSD_BWD(KM, 1, true); else MVSD_BWD(KM, 1, false); }        \
  } while (0)
  if (kmax == 1) MVSD_BWD_G(1);
  else if (kmax == 2) MVSD_BWD_G(2);
  else MVSD_BWD_G(4);
#undef MVSD_BWD_G
#undef MVSD_BWD
  count_launch();
  return check_launch("plane_sweep_bwd");
}

}  // namespace mvsd

using namespace mvsd;

extern "C" int mvsd_plane_sweep_bwd(const void* g_out, int g_dtype, int g_layout, const void* feat,
                                    int feat_dtype, const int32_t* nbr_ids, const float* hom,
                                    const float* depth_values, float* g_feat, int V, int C, int D,
                                    int H, int W, int k, int ref_begin, void* stream) {
  if (int e = sweep_check("plane_sweep_bwd", V, C, D, H, W, k, g_layout)) return e;
  if (!g_out || !feat || !g_feat || !depth_values || (k > 0 && (!nbr_ids || !hom)))
    return fail(MVSD_ERR_INVALID_ARG, "plane_sweep_bwd: null pointer");
  if (ref_begin < 0) return fail(MVSD_ERR_INVALID_ARG, "plane_sweep_bwd: negative ref_begin");
  SweepParams p{};
  p.feat = feat; p.nbr = nbr_ids; p.hom = hom; p.depth = depth_values; p.g_out = g_out;
  p.g_feat = g_feat;
  p.V = V; p.C = C; p.D = D; p.H = H; p.W = W; p.k = k; p.ref_begin = ref_begin;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (feat_dtype == MVSD_F32 && g_dtype == MVSD_F32) return launch_bwd_k<float, float, false>(p, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_F32)
    return launch_bwd_k<__nv_bfloat16, float, false>(p, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_BF16)
    return launch_bwd_k<__nv_bfloat16, __nv_bfloat16, false>(p, st);
  return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: dtype combination not built");
}

extern "C" int mvsd_homo_warp_bwd(const void* g_out, int g_dtype, int g_layout, const float* hom,
                                  const float* depth_values, float* g_src, int B, int C, int D,
                                  int H, int W, void* stream) {
  if (int e = sweep_check("homo_warp_bwd", B, C, D, H, W, 1, g_layout)) return e;
  if (!g_out || !hom || !depth_values || !g_src)
    return fail(MVSD_ERR_INVALID_ARG, "homo_warp_bwd: null pointer");
  SweepParams p{};
  p.feat = nullptr; p.nbr = nullptr; p.hom = hom; p.depth = depth_values; p.g_out = g_out;
  p.g_feat = g_src;
  p.V = B; p.C = C; p.D = D; p.H = H; p.W = W; p.k = 1; p.ref_begin = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_dtype == MVSD_F32) return launch_bwd_k<float, float, true>(p, st);
  if (g_dtype == MVSD_BF16) return launch_bwd_k<float, __nv_bfloat16, true>(p, st);
  return fail(MVSD_ERR_INVALID_ARG, "homo_warp_bwd: bad dtype");
}
