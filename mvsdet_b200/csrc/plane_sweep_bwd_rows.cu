// Plane-sweep backward, row-block variant (tuning key 5 = 5: two rows, 6: four rows).
//
// The run-merging kernel (plane_sweep_bwd_run.cu) is bound by the fp32 RED stream
// into L2 (3.79 GB of payload per scene against a measured 5.7 TB/s ceiling,
// DESIGN.md section 5): a warp walks ONE image row and merges only a pending right
// tap column with the next pixel's left column.  A host-side replay of that merge
// logic on the benchmark scene (tools/red_merge_sim.py) gives 3.69 GB; the same
// replay for the scheme below gives 2.73 GB (two rows) / 2.43 GB (four rows):
//
//   * a warp owns R consecutive rows x 8 columns x 128 channels (lane = 4 channels)
//     and walks the block column by column, planes in the outer loop;
//   * the scatter targets of one column are R+1 "sides" (source rows): the bottom
//     taps of row r and the top taps of row r+1 usually hit the same source row;
//   * every side keeps TWO pending source pixels (left, right) in registers.  A new
//     contribution (l, r) either lands on both (source x did not advance), shifts
//     the window by one (the usual case: one RED leaves), shifts it backwards, or
//     replaces it.  This also catches the +-1 column skew between rows of a rotated
//     neighbour that a single pending column misses.
//   * 128-channel warps keep the pending file at (R+1) x k x 2 x 4 registers.
// The per-pixel reference gradient (a sum over planes) lives in tensor memory
// (R*8*4 columns per CTA), as in the run kernel.
#include "plane_sweep.cuh"

namespace mvsd {
namespace {

constexpr int kRowCols = 8;                // pixels per block row
constexpr int kRowWarps = 4;               // warps (= row groups) per CTA
constexpr int kRowThreads = kRowWarps * 32;
constexpr unsigned kNone = 0xfffffffeu;    // empty pending slot / tap without weight

__device__ __forceinline__ void red_p4(float* p, P4 v) {
  float a, b, c, d;
  upk2(v.lo, a, b);
  upk2(v.hi, c, d);
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void pf_l2(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

struct Side {                              // two pending source pixels of one source row
  unsigned idl, idr;
  P4 al, ar;
};

__device__ __forceinline__ void side_flush_slot(float* dst, bool on, unsigned& id, const P4& acc) {
  if (id != kNone && on) red_p4(at(dst, id), acc);
  id = kNone;
}

// Adds gw*wl at source pixel pl and gw*wr at pr (a left/right tap pair of one source row).
__device__ __forceinline__ void side_add(float* dst, bool on, Side& s, const P4& gw, unsigned pl,
                                         float wl, unsigned pr, float wr) {
  const unsigned l = wl != 0.f ? pl : kNone;
  const unsigned r = wr != 0.f ? pr : kNone;
  const u64 wl2 = pk2(wl, wl), wr2 = pk2(wr, wr);
  if (l == s.idl && r == s.idr) {                       // same window
    s.al = p4fma(gw, wl2, s.al);
    s.ar = p4fma(gw, wr2, s.ar);
  } else if (l != kNone && l == s.idr) {                // window advanced by one
    side_flush_slot(dst, on, s.idl, s.al);
    s.al = p4fma(gw, wl2, s.ar);
    s.idl = l;
    s.ar = p4scale(gw, wr2);
    s.idr = r;
  } else if (r != kNone && r == s.idl) {                // window moved back by one
    side_flush_slot(dst, on, s.idr, s.ar);
    s.ar = p4fma(gw, wr2, s.al);
    s.idr = r;
    s.al = p4scale(gw, wl2);
    s.idl = l;
  } else {                                              // unrelated window
    side_flush_slot(dst, on, s.idl, s.al);
    side_flush_slot(dst, on, s.idr, s.ar);
    s.al = p4scale(gw, wl2);
    s.idl = l;
    s.ar = p4scale(gw, wr2);
    s.idr = r;
  }
}

template <typename TIn, typename TG, int KMAX, int R, bool FULL, int MINB>
__global__ void __launch_bounds__(kRowThreads, MINB) sweep_bwd_rows_kernel(const SweepParams p) {
  static_assert(R * kRowCols <= 32, "one prefetch lane per block pixel");
  constexpr int kTmemCols = R * kRowCols * 4;
  static_assert(kTmemCols >= 32 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns: power of two >= 32");
  constexpr int kTab = R * kRowCols * KMAX;
  __shared__ WarpSample s_tab[kRowWarps][kTab];
  __shared__ uint32_t s_tmem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x;
  const int xr = t % p.tiles_x; t /= p.tiles_x;
  const int yt = t % p.tiles_y; t /= p.tiles_y;
  const int slice = t % p.slices;
  const int v = t / p.slices;
  const int y0 = (yt * kRowWarps + warp) * R;
  const int x0 = xr * kRowCols;
  const int npix = min(kRowCols, p.W - x0);
  const int nrows = max(0, min(R, p.H - y0));
  const int c0 = slice * 128 + 4 * lane;
  const bool on = FULL || c0 < p.C;
  const uint32_t tbase = tmem_alloc_cta<kTmemCols>(&s_tmem, warp);
  if (nrows > 0) {
    const int C = p.C, k = p.k, W = p.W, HW = p.H * p.W;
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t ref_off = ((size_t)(v + p.ref_begin) * HW + (size_t)y0 * W + x0) * C + c0;
    const TIn* ref_blk = feat + ref_off;
    const size_t plane_stride = (size_t)HW * C;
    const TG* g_d = static_cast<const TG*>(p.g_out) + ((size_t)v * p.D * HW + (size_t)y0 * W + x0) * C + c0;
    WarpSample* tab = s_tab[warp];

    // L2 prefetch of the upstream gradient two planes ahead: lane <-> (row, column) of
    // the block, one 128-channel segment each.
    const unsigned pf_bytes = (unsigned)(min(128, C - slice * 128) * (int)sizeof(TG)) & ~15u;
    const int pf_r = lane / kRowCols, pf_i = lane % kRowCols;
    const TG* pf_base = g_d - 4 * lane + ((size_t)pf_r * W + pf_i) * C;
    const bool pf_ok = pf_bytes >= 16 && pf_r < nrows && pf_i < npix &&
                       (reinterpret_cast<uintptr_t>(pf_base) & 15) == 0 &&
                       ((plane_stride * sizeof(TG)) & 15) == 0;
    auto prefetch_plane = [&](int d) {
      if (!pf_ok || d >= p.D) return;
      pf_l2(pf_base + (size_t)d * plane_stride, pf_bytes);
    };
    prefetch_plane(0);
    prefetch_plane(1);

    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      int n = v + p.ref_begin;
      if (j < k) n = __ldg(p.nbr + (size_t)v * k + j);
      nsrc[j] = feat + (size_t)n * HW * C + c0;
      ndst[j] = p.g_feat + (size_t)n * HW * C + c0;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(k + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);

#pragma unroll
    for (int q = 0; q < R * kRowCols; ++q) tmem_st4(tbase + 4u * (uint32_t)q, p4zero());
    tmem_wait_st();

    for (int d = 0; d < p.D; ++d) {
      __syncwarp();
      // sample geometry of the block for this plane: entry (r*8 + i)*k + j
      for (int s = lane; s < R * kRowCols * k; s += 32) {
        const int r = s / (kRowCols * k), rem = s - r * (kRowCols * k);
        const int i = rem / k, j = rem - i * k;
        WarpSample ws;
        ws.w00 = ws.w01 = ws.w10 = ws.w11 = 0.f;
        ws.p00 = ws.p01 = ws.p10 = ws.p11 = kNoSample;
        if (r < nrows && i < npix) {
          const float* m = p.hom + ((size_t)v * k + j) * 12;
          float mm[12];
#pragma unroll
          for (int u = 0; u < 12; ++u) mm[u] = __ldg(m + u);
          ws = make_warp_sample(mm, (float)(x0 + i), (float)(y0 + r),
                                __ldg(p.depth + (size_t)v * p.D + d), p.H, W, C);
        }
        tab[s] = ws;
      }
      __syncwarp();
      prefetch_plane(d + 2);
      tmem_wait_st();

      Side side[R + 1][KMAX];
#pragma unroll
      for (int q = 0; q <= R; ++q)
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          side[q][j].idl = side[q][j].idr = kNone;
          side[q][j].al = side[q][j].ar = p4zero();
        }

#pragma unroll 1
      for (int i = 0; i < npix; ++i) {
        // every load of the column (R pixels: gradient, reference, k x 4 taps) goes out
        // before the first dependent instruction
        typename Raw<TG>::type graw[R];
        typename Raw<TIn>::type rraw[R];
        RawTaps<TIn, 1> traw[R][KMAX];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r >= nrows) continue;
          const size_t po = ((size_t)r * W + i) * C;
          graw[r] = on ? Raw<TG>::ld_stream_na(g_d + po) : Raw<TG>::zero();
          rraw[r] = on ? Raw<TIn>::ld(ref_blk + po) : Raw<TIn>::zero();
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            if (j >= k) continue;
            const WarpSample& ws = tab[(r * kRowCols + i) * k + j];
            if (ws.p00 != kNoSample) load_taps<TIn, 1, FULL>(nsrc[j], ws, c0, C, traw[r][j]);
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r >= nrows) continue;
          const P4 ref = p4from(rraw[r]);
          P4 mu = ref;
          P4 wv[KMAX][1];
          WarpSample smp[KMAX];
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            wv[j][0] = p4zero();
            if (j >= k) continue;
            smp[j] = tab[(r * kRowCols + i) * k + j];
            if (smp[j].p00 != kNoSample) blend_taps<TIn, 1>(traw[r][j], smp[j], wv[j]);
            mu = p4add(mu, wv[j][0]);
          }
          mu = p4scale(mu, inv_n2);
          const P4 gv = p4scale(p4from(graw[r]), two_inv_n2);
          const uint32_t ta = tbase + 4u * (uint32_t)(r * kRowCols + i);
          tmem_st4(ta, p4fma(gv, p4sub(ref, mu), tmem_ld4(ta)));
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            if (j >= k) continue;
            if (smp[j].p00 == kNoSample) continue;
            const P4 gw = p4mul(gv, p4sub(wv[j][0], mu));
            side_add(ndst[j], on, side[r][j], gw, smp[j].p00, smp[j].w00, smp[j].p01, smp[j].w01);
            side_add(ndst[j], on, side[r + 1][j], gw, smp[j].p10, smp[j].w10, smp[j].p11, smp[j].w11);
          }
        }
      }
#pragma unroll
      for (int q = 0; q <= R; ++q)
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j >= k) continue;
          side_flush_slot(ndst[j], on, side[q][j].idl, side[q][j].al);
          side_flush_slot(ndst[j], on, side[q][j].idr, side[q][j].ar);
        }
      g_d += plane_stride;
    }
    tmem_wait_st();
    float* dst = p.g_feat + ref_off;
    for (int r = 0; r < nrows; ++r)
      for (int i = 0; i < npix; ++i) {
        const P4 acc = tmem_ld4(tbase + 4u * (uint32_t)(r * kRowCols + i));
        if (on) red_p4(dst + ((size_t)r * W + i) * C, acc);
      }
  }
  tmem_free_cta<kTmemCols>(&s_tmem, warp);
}

template <typename TIn, typename TG>
int launch_rows_t(SweepParams& p, int rows, cudaStream_t st) {
  p.tiles_x = (p.W + kRowCols - 1) / kRowCols;
  p.tiles_y = (p.H + kRowWarps * rows - 1) / (kRowWarps * rows);
  p.slices = (p.C + 127) / 128;
  const long long blocks = (long long)p.V * p.slices * p.tiles_y * p.tiles_x;
  if (blocks > 2147483647LL) return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: grid too large");
  dim3 grid((unsigned)blocks);
  const bool full = p.C % 128 == 0;
#define MVSD_ROWS(KM, RR, FU, MB) \
  sweep_bwd_rows_kernel<TIn, TG, KM, RR, FU, MB><<<grid, kRowThreads, 0, st>>>(p)
#define MVSD_ROWS_K(RR, MB)                                                      \
  do {                                                                           \
    if (p.k == 1) { if (full) MVSD_ROWS(1, RR, true, MB); else MVSD_ROWS(1, RR, false, MB); } \
    else { if (full) MVSD_ROWS(2, RR, true, MB); else MVSD_ROWS(2, RR, false, MB); }          \
  } while (0)
  if (rows == 4) MVSD_ROWS_K(4, 2);
  else MVSD_ROWS_K(2, 3);
#undef MVSD_ROWS_K
#undef MVSD_ROWS
  count_launch();
  return check_launch("plane_sweep_bwd(rows)");
}

}  // namespace

int launch_bwd_rows(SweepParams& p, int rows, int feat_dtype, int g_dtype, cudaStream_t st) {
  if (feat_dtype == MVSD_F32 && g_dtype == MVSD_F32) return launch_rows_t<float, float>(p, rows, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_F32)
    return launch_rows_t<__nv_bfloat16, float>(p, rows, st);
  if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_BF16)
    return launch_rows_t<__nv_bfloat16, __nv_bfloat16>(p, rows, st);
  return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: dtype combination not built");
}

}  // namespace mvsd
