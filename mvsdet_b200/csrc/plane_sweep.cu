// Fused plane-sweep variance volume (forward + backward) and the stand-alone
// homography warp, for sm_100a.
//
// Replaces projects/NeRF-Det/nerfdet/mvsdet.py:439-467 and
// mvs_models/module.py:105-146 of the reference: instead of materialising the
// repeated reference volume, k warped [V,C,D,H,W] volumes, their squares and
// running sums (~30 passes over 1.18 GB tensors), one kernel gathers the
// bilinear taps of the k neighbours and writes the variance once.
//
// Mapping.  Features are channels-last ([V,H,W,C]), so one bilinear tap is a
// contiguous C-vector.  A warp owns a pixel: lane l holds channels
// [128 g + 4 l, 128 g + 4 l + 4) for g < G (16-byte vector loads, coalesced
// 512-byte requests).  The sample geometry (homography, divide, floor, tap
// weights) is warp-uniform, so it is computed ONCE per (pixel, plane,
// neighbour) by a single lane -- lane s handles sample slot s -- parked in
// shared memory and re-read by all lanes as a broadcast; the inner loop is then
// nothing but tap loads and FMAs.  A CTA of 8 warps covers a PW x PH pixel
// patch and walks the depth planes in the outer loop, so at any moment the CTA
// touches a compact source footprint that stays in L1 (bilinear overlap
// between neighbouring pixels is served there, not from L2).
//
// HBM-bound by design: no tensor cores (there is no contraction here),
// SURVEY.md section 0.  Algorithmic bytes / scene are in DESIGN.md.
#include "common.cuh"

namespace mvsd {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSlots = 32;   // sample slots per pixel per chunk (one per lane)

struct SweepParams {
  const void* feat;        // nhwc features (fwd, bwd variance)
  const int32_t* nbr;      // [V,k] or nullptr (warp-only: source = same index)
  const float* hom;        // [V,k,12]
  const float* depth;      // [V,D]
  void* out;               // fwd output [V,D,H,W,C]
  const void* g_out;       // bwd upstream gradient, same layout
  float* g_feat;           // bwd: nhwc fp32, accumulated with RED
  int V, C, D, H, W, k;
  int ref_begin;           // feat index of reference view 0 (view sharding)
  int pw, ph;              // patch width / height in pixels, pw*ph == kWarps*PPW
  int tiles_x, tiles_y;
};

// Fill this warp's sample table for planes [d0, d0+dc) of pixel slot i.
__device__ __forceinline__ void fill_samples(WarpSample* tab, const SweepParams& p, int v,
                                             int x, int y, int d0, int dc, int lane) {
  const int k = p.k;
  const int nslot = dc * k;
  if (lane < nslot) {
    const int dd = lane / k, j = lane - dd * k;
    const int d = d0 + dd;
    WarpSample s;
    if (d < p.D) {
      const float* m = p.hom + ((size_t)v * k + j) * 12;
      float mm[12];
#pragma unroll
      for (int t = 0; t < 12; ++t) mm[t] = __ldg(m + t);
      s = make_warp_sample(mm, (float)x, (float)y, __ldg(p.depth + (size_t)v * p.D + d), p.H, p.W);
    } else {
      s.w00 = s.w01 = s.w10 = s.w11 = 0.f;
      s.p00 = s.p01 = s.p10 = s.p11 = -1;
    }
    tab[lane] = s;
  }
}

template <typename TIn, int G>
__device__ __forceinline__ void gather_bilinear(const TIn* __restrict__ src, const WarpSample& s,
                                                int C, int lane, float4 (&wv)[G]) {
  const TIn* b00 = src + (size_t)s.p00 * C + 4 * lane;
  const TIn* b01 = src + (size_t)s.p01 * C + 4 * lane;
  const TIn* b10 = src + (size_t)s.p10 * C + 4 * lane;
  const TIn* b11 = src + (size_t)s.p11 * C + 4 * lane;
  float4 a[G], b[G], c[G], e[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (128 * g + 4 * lane < C) {
      a[g] = Io<TIn>::ld(b00 + 128 * g);
      b[g] = Io<TIn>::ld(b01 + 128 * g);
      c[g] = Io<TIn>::ld(b10 + 128 * g);
      e[g] = Io<TIn>::ld(b11 + 128 * g);
    } else {
      a[g] = b[g] = c[g] = e[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    // (nw*w00 + ne*w01) + sw*w10 + se*w11, the order ATen's sampler uses
    wv[g].x = fmaf(e[g].x, s.w11, fmaf(c[g].x, s.w10, fmaf(b[g].x, s.w01, a[g].x * s.w00)));
    wv[g].y = fmaf(e[g].y, s.w11, fmaf(c[g].y, s.w10, fmaf(b[g].y, s.w01, a[g].y * s.w00)));
    wv[g].z = fmaf(e[g].z, s.w11, fmaf(c[g].z, s.w10, fmaf(b[g].z, s.w01, a[g].z * s.w00)));
    wv[g].w = fmaf(e[g].w, s.w11, fmaf(c[g].w, s.w10, fmaf(b[g].w, s.w01, a[g].w * s.w00)));
  }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <typename TIn, typename TOut, int G, int PPW, bool WARP_ONLY>
__global__ void __launch_bounds__(kThreads) sweep_fwd_kernel(const SweepParams p) {
  __shared__ WarpSample s_tab[kWarps][PPW][kSlots];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int v = t / p.tiles_y;
  const int HW = p.H * p.W, C = p.C, k = p.k;
  const TIn* feat = static_cast<const TIn*>(p.feat);
  TOut* out = static_cast<TOut*>(p.out);

  int px[PPW], py[PPW];
  bool pok[PPW];
  float4 ref[PPW][G];
#pragma unroll
  for (int i = 0; i < PPW; ++i) {
    const int q = i * kWarps + warp;
    px[i] = tx * p.pw + q % p.pw;
    py[i] = ty * p.ph + q / p.pw;
    pok[i] = px[i] < p.W && py[i] < p.H;
#pragma unroll
    for (int g = 0; g < G; ++g) ref[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!WARP_ONLY && pok[i]) {
      const TIn* r = feat + ((size_t)(v + p.ref_begin) * HW + (size_t)py[i] * p.W + px[i]) * C + 4 * lane;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (128 * g + 4 * lane < C) ref[i][g] = Io<TIn>::ld(r + 128 * g);
    }
  }
  // neighbour map base pointers (warp-uniform)
  const TIn* nsrc[MVSD_MAX_K];
#pragma unroll
  for (int j = 0; j < MVSD_MAX_K; ++j) {
    int n = v + p.ref_begin;
    if (!WARP_ONLY && j < k) n = __ldg(p.nbr + (size_t)v * k + j);
    nsrc[j] = feat + (size_t)n * HW * C;
  }
  const float inv_n = 1.0f / (float)(k + 1);
  const int kk = k > 0 ? k : 1;
  const int dc = kSlots / kk;                 // planes per chunk

  for (int d0 = 0; d0 < p.D; d0 += dc) {
    if (k > 0) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < PPW; ++i)
        if (pok[i]) fill_samples(s_tab[warp][i], p, v, px[i], py[i], d0, dc, lane);
      __syncwarp();
    }
    const int dend = min(p.D, d0 + dc);
    for (int d = d0; d < dend; ++d) {
#pragma unroll
      for (int i = 0; i < PPW; ++i) {
        if (!pok[i]) continue;
        float4 s1[G], s2[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
          s1[g] = ref[i][g];
          s2[g] = make_float4(ref[i][g].x * ref[i][g].x, ref[i][g].y * ref[i][g].y,
                              ref[i][g].z * ref[i][g].z, ref[i][g].w * ref[i][g].w);
        }
#pragma unroll
        for (int j = 0; j < MVSD_MAX_K; ++j) {
          if (j >= k) break;
          const WarpSample s = s_tab[warp][i][(d - d0) * k + j];
          if (s.p00 < 0) continue;            // all four taps fall outside: adds 0
          float4 wv[G];
          gather_bilinear<TIn, G>(nsrc[j], s, C, lane, wv);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            s1[g].x += wv[g].x; s1[g].y += wv[g].y; s1[g].z += wv[g].z; s1[g].w += wv[g].w;
            s2[g].x = fmaf(wv[g].x, wv[g].x, s2[g].x);
            s2[g].y = fmaf(wv[g].y, wv[g].y, s2[g].y);
            s2[g].z = fmaf(wv[g].z, wv[g].z, s2[g].z);
            s2[g].w = fmaf(wv[g].w, wv[g].w, s2[g].w);
          }
        }
        TOut* o = out + (((size_t)v * p.D + d) * HW + (size_t)py[i] * p.W + px[i]) * C + 4 * lane;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (128 * g + 4 * lane >= C) continue;
          float4 r;
          if (WARP_ONLY) {
            r = s1[g];
          } else {
            // var = S2/n - (S1/n)^2  (mvsdet.py:467); /n as *(1/n), which is what
            // ATen's CUDA div-by-scalar does
            const float mx = s1[g].x * inv_n, my = s1[g].y * inv_n;
            const float mz = s1[g].z * inv_n, mw = s1[g].w * inv_n;
            r.x = __fsub_rn(s2[g].x * inv_n, __fmul_rn(mx, mx));
            r.y = __fsub_rn(s2[g].y * inv_n, __fmul_rn(my, my));
            r.z = __fsub_rn(s2[g].z * inv_n, __fmul_rn(mz, mz));
            r.w = __fsub_rn(s2[g].w * inv_n, __fmul_rn(mw, mw));
          }
          Io<TOut>::st_stream(o + 128 * g, r);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// backward
//   var = S2/n - mu^2, mu = S1/n
//   dL/dref    += sum_d G (2/n) (ref - mu_d)
//   dL/dwarp_j  = G (2/n) (warp_j - mu)   -> bilinear scatter into neighbour j
// (SURVEY.md Appendix A.4).  Warp-only mode: dL/dsrc = scatter of G.
// ---------------------------------------------------------------------------
template <typename TIn, typename TG, int G, int PPW, int KMAX, bool WARP_ONLY>
__global__ void __launch_bounds__(kThreads) sweep_bwd_kernel(const SweepParams p) {
  __shared__ WarpSample s_tab[kWarps][PPW][kSlots];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int v = t / p.tiles_y;
  const int HW = p.H * p.W, C = p.C, k = p.k;
  const TIn* feat = static_cast<const TIn*>(p.feat);
  const TG* gout = static_cast<const TG*>(p.g_out);

  int px[PPW], py[PPW];
  bool pok[PPW];
  float4 ref[PPW][G], gref[PPW][G];
#pragma unroll
  for (int i = 0; i < PPW; ++i) {
    const int q = i * kWarps + warp;
    px[i] = tx * p.pw + q % p.pw;
    py[i] = ty * p.ph + q / p.pw;
    pok[i] = px[i] < p.W && py[i] < p.H;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      ref[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
      gref[i][g] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!WARP_ONLY && pok[i]) {
      const TIn* r = feat + ((size_t)(v + p.ref_begin) * HW + (size_t)py[i] * p.W + px[i]) * C + 4 * lane;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (128 * g + 4 * lane < C) ref[i][g] = Io<TIn>::ld(r + 128 * g);
    }
  }
  int nidx[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    nidx[j] = v + p.ref_begin;
    if (!WARP_ONLY && j < k) nidx[j] = __ldg(p.nbr + (size_t)v * k + j);
  }
  const float inv_n = 1.0f / (float)(k + 1);
  const float two_inv_n = 2.0f * inv_n;
  const int kk = k > 0 ? k : 1;
  const int dc = kSlots / kk;

  for (int d0 = 0; d0 < p.D; d0 += dc) {
    if (k > 0) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < PPW; ++i)
        if (pok[i]) fill_samples(s_tab[warp][i], p, v, px[i], py[i], d0, dc, lane);
      __syncwarp();
    }
    const int dend = min(p.D, d0 + dc);
    for (int d = d0; d < dend; ++d) {
#pragma unroll
      for (int i = 0; i < PPW; ++i) {
        if (!pok[i]) continue;
        const TG* gp = gout + (((size_t)v * p.D + d) * HW + (size_t)py[i] * p.W + px[i]) * C + 4 * lane;
        float4 gv[G];
#pragma unroll
        for (int g = 0; g < G; ++g)
          gv[g] = (128 * g + 4 * lane < C) ? Io<TG>::ld_stream(gp + 128 * g)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 wv[KMAX][G];
        float4 mu[G];
        if (!WARP_ONLY) {
#pragma unroll
          for (int g = 0; g < G; ++g) mu[g] = ref[i][g];
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
#pragma unroll
            for (int g = 0; g < G; ++g) wv[j][g] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j >= k) continue;
            const WarpSample s = s_tab[warp][i][(d - d0) * k + j];
            if (s.p00 >= 0) gather_bilinear<TIn, G>(feat + (size_t)nidx[j] * HW * C, s, C, lane, wv[j]);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              mu[g].x += wv[j][g].x; mu[g].y += wv[j][g].y;
              mu[g].z += wv[j][g].z; mu[g].w += wv[j][g].w;
            }
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            mu[g].x *= inv_n; mu[g].y *= inv_n; mu[g].z *= inv_n; mu[g].w *= inv_n;
            gv[g].x *= two_inv_n; gv[g].y *= two_inv_n; gv[g].z *= two_inv_n; gv[g].w *= two_inv_n;
            gref[i][g].x = fmaf(gv[g].x, ref[i][g].x - mu[g].x, gref[i][g].x);
            gref[i][g].y = fmaf(gv[g].y, ref[i][g].y - mu[g].y, gref[i][g].y);
            gref[i][g].z = fmaf(gv[g].z, ref[i][g].z - mu[g].z, gref[i][g].z);
            gref[i][g].w = fmaf(gv[g].w, ref[i][g].w - mu[g].w, gref[i][g].w);
          }
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j >= k) continue;
          const WarpSample s = s_tab[warp][i][(d - d0) * k + j];
          if (s.p00 < 0) continue;
          float* dst = p.g_feat + (size_t)nidx[j] * HW * C + 4 * lane;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (128 * g + 4 * lane >= C) continue;
            float4 gw;
            if (WARP_ONLY) {
              gw = gv[g];
            } else {
              gw.x = gv[g].x * (wv[j][g].x - mu[g].x);
              gw.y = gv[g].y * (wv[j][g].y - mu[g].y);
              gw.z = gv[g].z * (wv[j][g].z - mu[g].z);
              gw.w = gv[g].w * (wv[j][g].w - mu[g].w);
            }
            if (s.w00 != 0.f)
              red_add_f32x4(dst + (size_t)s.p00 * C + 128 * g,
                            make_float4(gw.x * s.w00, gw.y * s.w00, gw.z * s.w00, gw.w * s.w00));
            if (s.w01 != 0.f)
              red_add_f32x4(dst + (size_t)s.p01 * C + 128 * g,
                            make_float4(gw.x * s.w01, gw.y * s.w01, gw.z * s.w01, gw.w * s.w01));
            if (s.w10 != 0.f)
              red_add_f32x4(dst + (size_t)s.p10 * C + 128 * g,
                            make_float4(gw.x * s.w10, gw.y * s.w10, gw.z * s.w10, gw.w * s.w10));
            if (s.w11 != 0.f)
              red_add_f32x4(dst + (size_t)s.p11 * C + 128 * g,
                            make_float4(gw.x * s.w11, gw.y * s.w11, gw.z * s.w11, gw.w * s.w11));
          }
        }
      }
    }
  }
  if (!WARP_ONLY) {
#pragma unroll
    for (int i = 0; i < PPW; ++i) {
      if (!pok[i]) continue;
      float* dst = p.g_feat + ((size_t)(v + p.ref_begin) * HW + (size_t)py[i] * p.W + px[i]) * C + 4 * lane;
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (128 * g + 4 * lane < C) red_add_f32x4(dst + 128 * g, gref[i][g]);
    }
  }
}

// ---------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------
static bool pick_patch(int ppw, int& pw, int& ph) {
  const int npix = kWarps * ppw;
  pw = tuning(1);
  if (pw <= 0 || npix % pw != 0) pw = npix >= 32 ? 8 : (npix >= 16 ? 4 : 4);
  if (npix % pw != 0) return false;
  ph = npix / pw;
  return true;
}

template <typename TIn, typename TOut, int G, bool WARP_ONLY>
static int launch_fwd_ppw(SweepParams& p, int ppw, cudaStream_t st) {
  if (!pick_patch(ppw, p.pw, p.ph)) return fail(MVSD_ERR_INVALID_ARG, "bad patch shape");
  p.tiles_x = (p.W + p.pw - 1) / p.pw;
  p.tiles_y = (p.H + p.ph - 1) / p.ph;
  const long long blocks = (long long)p.V * p.tiles_x * p.tiles_y;
  if (blocks > 2147483647LL) return fail(MVSD_ERR_UNSUPPORTED, "grid too large");
  dim3 grid((unsigned)blocks);
  switch (ppw) {
    case 1: sweep_fwd_kernel<TIn, TOut, G, 1, WARP_ONLY><<<grid, kThreads, 0, st>>>(p); break;
    case 2: sweep_fwd_kernel<TIn, TOut, G, 2, WARP_ONLY><<<grid, kThreads, 0, st>>>(p); break;
    default: sweep_fwd_kernel<TIn, TOut, G, 4, WARP_ONLY><<<grid, kThreads, 0, st>>>(p); break;
  }
  count_launch();
  return check_launch("plane_sweep_fwd");
}

template <typename TIn, typename TOut, bool WARP_ONLY>
static int launch_fwd_g(SweepParams& p, cudaStream_t st) {
  int ppw = tuning(0);
  if (ppw != 1 && ppw != 2 && ppw != 4) ppw = 4;
  const int G = (p.C + 127) / 128;
  switch (G) {
    case 1: return launch_fwd_ppw<TIn, TOut, 1, WARP_ONLY>(p, ppw, st);
    case 2: return launch_fwd_ppw<TIn, TOut, 2, WARP_ONLY>(p, ppw, st);
    default: return launch_fwd_ppw<TIn, TOut, 4, WARP_ONLY>(p, ppw, st);
  }
}

template <bool WARP_ONLY>
static int launch_fwd(SweepParams& p, int in_dtype, int out_dtype, cudaStream_t st) {
  if (in_dtype == MVSD_F32 && out_dtype == MVSD_F32)
    return launch_fwd_g<float, float, WARP_ONLY>(p, st);
  if (in_dtype == MVSD_BF16 && out_dtype == MVSD_F32)
    return launch_fwd_g<__nv_bfloat16, float, WARP_ONLY>(p, st);
  if (in_dtype == MVSD_BF16 && out_dtype == MVSD_BF16)
    return launch_fwd_g<__nv_bfloat16, __nv_bfloat16, WARP_ONLY>(p, st);
  if (in_dtype == MVSD_F32 && out_dtype == MVSD_BF16)
    return launch_fwd_g<float, __nv_bfloat16, WARP_ONLY>(p, st);
  return fail(MVSD_ERR_INVALID_ARG, "bad dtype");
}

template <typename TIn, typename TG, int G, bool WARP_ONLY>
static int launch_bwd_k(SweepParams& p, cudaStream_t st) {
  int ppw = tuning(2);
  if (ppw != 1 && ppw != 2) ppw = 2;
  if (!pick_patch(ppw, p.pw, p.ph)) return fail(MVSD_ERR_INVALID_ARG, "bad patch shape");
  p.tiles_x = (p.W + p.pw - 1) / p.pw;
  p.tiles_y = (p.H + p.ph - 1) / p.ph;
  const long long blocks = (long long)p.V * p.tiles_x * p.tiles_y;
  if (blocks > 2147483647LL) return fail(MVSD_ERR_UNSUPPORTED, "grid too large");
  dim3 grid((unsigned)blocks);
  if (p.k <= 2) {
    if (ppw == 1) sweep_bwd_kernel<TIn, TG, G, 1, 2, WARP_ONLY><<<grid, kThreads, 0, st>>>(p);
    else sweep_bwd_kernel<TIn, TG, G, 2, 2, WARP_ONLY><<<grid, kThreads, 0, st>>>(p);
  } else {
    if (ppw == 1) sweep_bwd_kernel<TIn, TG, G, 1, MVSD_MAX_K, WARP_ONLY><<<grid, kThreads, 0, st>>>(p);
    else sweep_bwd_kernel<TIn, TG, G, 2, MVSD_MAX_K, WARP_ONLY><<<grid, kThreads, 0, st>>>(p);
  }
  count_launch();
  return check_launch("plane_sweep_bwd");
}

template <typename TIn, typename TG, bool WARP_ONLY>
static int launch_bwd_g(SweepParams& p, cudaStream_t st) {
  const int G = (p.C + 127) / 128;
  switch (G) {
    case 1: return launch_bwd_k<TIn, TG, 1, WARP_ONLY>(p, st);
    case 2: return launch_bwd_k<TIn, TG, 2, WARP_ONLY>(p, st);
    default: return launch_bwd_k<TIn, TG, 4, WARP_ONLY>(p, st);
  }
}

static int check_common(const char* who, int V, int C, int D, int H, int W, int k, int layout) {
  if (V <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0 || k < 0)
    return fail(MVSD_ERR_INVALID_ARG, "%s: non-positive dimension", who);
  if (C % 4 != 0 || C > MVSD_MAX_C)
    return fail(MVSD_ERR_UNSUPPORTED, "%s: C=%d must be a multiple of 4 and <= %d", who, C, MVSD_MAX_C);
  if (k > MVSD_MAX_K) return fail(MVSD_ERR_UNSUPPORTED, "%s: k=%d > %d", who, k, MVSD_MAX_K);
  if ((long long)H * W >= (1LL << 30)) return fail(MVSD_ERR_UNSUPPORTED, "%s: map too large", who);
  if (layout != MVSD_CHANNELS_LAST)
    return fail(MVSD_ERR_UNSUPPORTED, "%s: only MVSD_CHANNELS_LAST volumes are implemented", who);
  return MVSD_OK;
}

}  // namespace mvsd

using namespace mvsd;

extern "C" int mvsd_plane_sweep_fwd(const void* feat, int feat_dtype, const int32_t* nbr_ids,
                                    const float* hom, const float* depth_values, void* out,
                                    int out_dtype, int out_layout, int V, int C, int D, int H,
                                    int W, int k, int ref_begin, void* stream) {
  if (int e = check_common("plane_sweep_fwd", V, C, D, H, W, k, out_layout)) return e;
  if (!feat || !out || !depth_values || (k > 0 && (!nbr_ids || !hom)))
    return fail(MVSD_ERR_INVALID_ARG, "plane_sweep_fwd: null pointer");
  SweepParams p{};
  p.feat = feat; p.nbr = nbr_ids; p.hom = hom; p.depth = depth_values; p.out = out;
  p.V = V; p.C = C; p.D = D; p.H = H; p.W = W; p.k = k; p.ref_begin = ref_begin;
  if (ref_begin < 0) return fail(MVSD_ERR_INVALID_ARG, "plane_sweep_fwd: negative ref_begin");
  return launch_fwd<false>(p, feat_dtype, out_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int mvsd_plane_sweep_bwd(const void* g_out, int g_dtype, int g_layout, const void* feat,
                                    int feat_dtype, const int32_t* nbr_ids, const float* hom,
                                    const float* depth_values, float* g_feat, int V, int C, int D,
                                    int H, int W, int k, int ref_begin, void* stream) {
  if (int e = check_common("plane_sweep_bwd", V, C, D, H, W, k, g_layout)) return e;
  if (!g_out || !feat || !g_feat || !depth_values || (k > 0 && (!nbr_ids || !hom)))
    return fail(MVSD_ERR_INVALID_ARG, "plane_sweep_bwd: null pointer");
  SweepParams p{};
  p.feat = feat; p.nbr = nbr_ids; p.hom = hom; p.depth = depth_values; p.g_out = g_out;
  p.g_feat = g_feat;
  p.V = V; p.C = C; p.D = D; p.H = H; p.W = W; p.k = k; p.ref_begin = ref_begin;
  if (ref_begin < 0) return fail(MVSD_ERR_INVALID_ARG, "plane_sweep_bwd: negative ref_begin");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (feat_dtype == MVSD_F32 && g_dtype == MVSD_F32) return launch_bwd_g<float, float, false>(p, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_F32)
    return launch_bwd_g<__nv_bfloat16, float, false>(p, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_BF16)
    return launch_bwd_g<__nv_bfloat16, __nv_bfloat16, false>(p, st);
  return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: dtype combination not built");
}

extern "C" int mvsd_homo_warp_fwd(const void* src, int src_dtype, const float* hom,
                                  const float* depth_values, void* out, int out_dtype,
                                  int out_layout, int B, int C, int D, int H, int W,
                                  void* stream) {
  if (int e = check_common("homo_warp_fwd", B, C, D, H, W, 1, out_layout)) return e;
  if (!src || !hom || !depth_values || !out)
    return fail(MVSD_ERR_INVALID_ARG, "homo_warp_fwd: null pointer");
  SweepParams p{};
  p.feat = src; p.nbr = nullptr; p.hom = hom; p.depth = depth_values; p.out = out;
  p.V = B; p.C = C; p.D = D; p.H = H; p.W = W; p.k = 1;
  return launch_fwd<true>(p, src_dtype, out_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int mvsd_homo_warp_bwd(const void* g_out, int g_dtype, int g_layout, const float* hom,
                                  const float* depth_values, float* g_src, int B, int C, int D,
                                  int H, int W, void* stream) {
  if (int e = check_common("homo_warp_bwd", B, C, D, H, W, 1, g_layout)) return e;
  if (!g_out || !hom || !depth_values || !g_src)
    return fail(MVSD_ERR_INVALID_ARG, "homo_warp_bwd: null pointer");
  SweepParams p{};
  p.feat = nullptr; p.nbr = nullptr; p.hom = hom; p.depth = depth_values; p.g_out = g_out;
  p.g_feat = g_src;
  p.V = B; p.C = C; p.D = D; p.H = H; p.W = W; p.k = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (g_dtype == MVSD_F32) return launch_bwd_g<float, float, true>(p, st);
  if (g_dtype == MVSD_BF16) return launch_bwd_g<float, __nv_bfloat16, true>(p, st);
  return fail(MVSD_ERR_INVALID_ARG, "homo_warp_bwd: bad dtype");
}
