// Plane-sweep backward, block-merging variant (default for k <= 2), sm_100a.
//
// Why.  tools/microbench_red.cu measures the chip-wide ceiling of fp32 reductions
// into L2 at ~5.7 TB/s of payload (the L2 atomic ALU retires one 32 B sector per
// 2 clk per slice; TMA bulk reductions hit the same ceiling).  The run-merging
// kernel (plane_sweep_bwd_run.cu) sends 3.8 GB of REDs per 20-view scene: 0.66 ms
// of its 0.86 ms is that ceiling.  The only way down is to merge more
// contributions on chip before they leave the SM.
//
// How.  A warp owns a block of kBlkRows x kRun reference pixels and 128 channels
// (lane = 4 channels) and walks it plane by plane, row by row:
//   * horizontally, the contribution to the right tap column stays pending in
//     registers and merges with the next pixel's left column (as before);
//   * vertically, finished BOTTOM-row cells are parked in a warp-private row
//     cache in shared memory (kSlots cells per neighbour, direct-mapped by source
//     column, tagged with the cell offset); when the next reference row finishes a
//     TOP-row cell it picks up the parked partner and leaves as ONE RED.  Tags make
//     this exact for any homography -- a miss only costs the merge.
//   * the per-pixel reference gradient (a sum over planes) needs 512 B per pixel
//     per warp; for a 4x8 block that is 16 KB per warp, which shared memory cannot
//     hold next to the row cache at 16 warps/SM.  It lives in TENSOR MEMORY: each
//     CTA allocates 128 TMEM columns, each warp uses its own 32-lane quarter as a
//     32-pixel x 4-float accumulator file (tcgen05.st / tcgen05.ld, 32x32b shape,
//     dynamic column index).  tools/tmem_probe.cu checks exactly this usage.
// Simulated on the bench scene: 1.83 REDs per valid sample against
// 2.7 for row-only merging.
#include "plane_sweep.cuh"

namespace mvsd {

constexpr int kBRun = 8;                   // pixels per row of a block
constexpr int kBRows = 4;                  // rows per block
constexpr int kBWarps = 4;                 // warps (blocks) per CTA
constexpr int kBThreads = kBWarps * 32;
constexpr int kBSlots = 12;                // row-cache cells per neighbour
constexpr int kBTmemCols = kBRun * kBRows * 4;   // 128: 4 fp32 columns per pixel
constexpr unsigned kNoCell = 0xfffffffeu;
static_assert(kBTmemCols == 128, "TMEM allocation must be a power of two >= 32 columns");

__device__ __forceinline__ void red_p4(float* p, P4 v) {
  float a, b, c, d;
  upk2(v.lo, a, b);
  upk2(v.hi, c, d);
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct BlkShared {
  WarpSample tab[kBWarps][2 * kBRun];               // samples of the current row (pixel-major, then neighbour)
  unsigned slots[kBWarps][2 * kBRun];               // row-cache slot of the left | right<<8 tap column
  P4 cache[kBWarps][2][kBSlots][32];                // parked bottom-row cells
  unsigned tag[kBWarps][2][kBSlots];                // their cell offsets (kNoCell = empty)
  uint32_t tmem_base;
};

// A finished TOP-row cell: pick up the parked partner (if any) and leave.
__device__ __forceinline__ void emit_top(bool on, BlkShared& sh, int warp, int lane, int j, float* dst,
                                         unsigned cell, unsigned slot, P4 v) {
  const unsigned t = sh.tag[warp][j][slot];
  if (t == cell) {
    const P4 c = sh.cache[warp][j][slot][lane];
    v = p4add(v, c);
    __syncwarp();
    if (lane == 0) sh.tag[warp][j][slot] = kNoCell;
    __syncwarp();
  }
  if (on) red_p4(at(dst, cell), v);
}

// A finished BOTTOM-row cell: park it for the next reference row.
__device__ __forceinline__ void emit_bot(bool on, BlkShared& sh, int warp, int lane, int j, float* dst,
                                         unsigned cell, unsigned slot, P4 v) {
  const unsigned t = sh.tag[warp][j][slot];
  if (t == cell) {
    v = p4add(v, sh.cache[warp][j][slot][lane]);
  } else if (t != kNoCell) {
    if (on) red_p4(at(dst, t), sh.cache[warp][j][slot][lane]);
  }
  sh.cache[warp][j][slot][lane] = v;
  __syncwarp();
  if (lane == 0) sh.tag[warp][j][slot] = cell;
  __syncwarp();
}

// One row (top or bottom) of the scatter of one sample; `TOP` picks the emitter.
template <bool TOP>
__device__ __forceinline__ void blk_side(bool on, BlkShared& sh, int warp, int lane, int j, float* dst,
                                         const P4& gw, float w_left, float w_right, unsigned p_left,
                                         unsigned p_right, unsigned s_left, unsigned s_right,
                                         unsigned& open_id, unsigned& open_slot, P4& open) {
  const u64 wl = pk2(w_left, w_left), wr = pk2(w_right, w_right);
  if (open_id == p_left) {
    const P4 a = p4fma(gw, wl, open);
    if (TOP) emit_top(on, sh, warp, lane, j, dst, p_left, s_left, a);
    else emit_bot(on, sh, warp, lane, j, dst, p_left, s_left, a);
  } else {
    if (open_id != kNoCell) {
      if (TOP) emit_top(on, sh, warp, lane, j, dst, open_id, open_slot, open);
      else emit_bot(on, sh, warp, lane, j, dst, open_id, open_slot, open);
    }
    if (w_left != 0.f) {
      const P4 a = p4scale(gw, wl);
      if (TOP) emit_top(on, sh, warp, lane, j, dst, p_left, s_left, a);
      else emit_bot(on, sh, warp, lane, j, dst, p_left, s_left, a);
    }
  }
  open = p4scale(gw, wr);
  open_id = w_right != 0.f ? p_right : kNoCell;
  open_slot = s_right;
}

template <bool TOP>
__device__ __forceinline__ void blk_flush(bool on, BlkShared& sh, int warp, int lane, int j, float* dst,
                                          unsigned& open_id, unsigned open_slot, const P4& open) {
  if (open_id != kNoCell) {
    if (TOP) emit_top(on, sh, warp, lane, j, dst, open_id, open_slot, open);
    else emit_bot(on, sh, warp, lane, j, dst, open_id, open_slot, open);
  }
  open_id = kNoCell;
}

template <typename TIn, typename TG, int KMAX, bool FULL>
__global__ void __launch_bounds__(kBThreads, 4) sweep_bwd_blk_kernel(const SweepParams p, int n_items,
                                                                    int n_xr, int n_rb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlkShared& sh = *reinterpret_cast<BlkShared*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- TMEM: 128 columns per CTA, one 32-lane quarter per warp ----------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&sh.tmem_base)), "n"(kBTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = sh.tmem_base + ((uint32_t)(warp * 32) << 16);

  const int item = blockIdx.x * kBWarps + warp;
  if (item < n_items) {
    int t = item;
    const int xr = t % n_xr; t /= n_xr;
    const int rb = t % n_rb; t /= n_rb;
    const int slice = t % p.slices;
    const int v = t / p.slices;
    const int C = p.C, k = p.k, HW = p.H * p.W;
    const int x0 = xr * kBRun, y0 = rb * kBRows;
    const int npix = min(kBRun, p.W - x0), nrows = min(kBRows, p.H - y0);
    const int c0 = slice * 128 + 4 * lane;
    const bool on = FULL || c0 < C;
    const int cc = on ? c0 : 0;                       // inactive lanes shadow channel 0 (never stored)
    const TIn* feat = static_cast<const TIn*>(p.feat);
    const size_t plane_stride = (size_t)HW * C;
    const size_t blk_off = ((size_t)y0 * p.W + x0) * C + cc;
    const TIn* ref_blk = feat + (size_t)(v + p.ref_begin) * HW * C + blk_off;
    const TG* g_blk = static_cast<const TG*>(p.g_out) + (size_t)v * p.D * plane_stride + blk_off;
    const TG* g_slice = static_cast<const TG*>(p.g_out) + (size_t)v * p.D * plane_stride +
                        ((size_t)y0 * p.W + x0) * C + slice * 128;
    const TIn* nsrc[KMAX];
    float* ndst[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      int n = v + p.ref_begin;
      if (j < k) n = __ldg(p.nbr + (size_t)v * k + j);
      nsrc[j] = feat + (size_t)n * HW * C + cc;
      ndst[j] = p.g_feat + (size_t)n * HW * C + cc;
      asm volatile("" : "+l"(nsrc[j]), "+l"(ndst[j]));
    }
    const float inv_n = 1.0f / (float)(k + 1);
    const u64 inv_n2 = pk2(inv_n, inv_n);
    const u64 two_inv_n2 = pk2(2.0f * inv_n, 2.0f * inv_n);

    for (int c = 0; c < kBTmemCols; c += 4) tmem_st4(tbase + c, p4zero());
    if (lane < 2 * kBSlots) sh.tag[warp][lane / kBSlots][lane % kBSlots] = kNoCell;
    tmem_wait_st();
    __syncwarp();

    // L2 prefetch of this warp's slice of the upstream gradient: 128 ch x 4 B = 512 B = 4 lines
    // per pixel; lane -> (pixel lane/4, line lane%4), one instruction per block row.
    const int slice_lines = (min(128, C - slice * 128) * (int)sizeof(TG) + 127) / 128;
    auto prefetch_plane = [&](int d) {
      if (d >= p.D) return;
      const int px = lane >> 2, ln = lane & 3;
      if (px < npix && ln < slice_lines) {
        const char* q = reinterpret_cast<const char*>(g_slice + (size_t)d * plane_stride + (size_t)px * C) + 128 * ln;
        for (int r = 0; r < nrows; ++r)
          asm volatile("prefetch.global.L2 [%0];" :: "l"(q + (size_t)r * p.W * C * sizeof(TG)));
      }
    };
    prefetch_plane(0);
    prefetch_plane(1);

    for (int d = 0; d < p.D; ++d) {
      prefetch_plane(d + 2);
      const float depth = __ldg(p.depth + (size_t)v * p.D + d);
      const TG* g_d = g_blk + (size_t)d * plane_stride;
      for (int r = 0; r < nrows; ++r) {
        // ---- sample geometry of this row: lane -> (pixel lane / k, neighbour lane % k)
        __syncwarp();
        if (lane < npix * k) {
          const int i = lane / k, j = lane - i * k;
          const float* m = p.hom + ((size_t)v * k + j) * 12;
          float mm[12];
#pragma unroll
          for (int q = 0; q < 12; ++q) mm[q] = __ldg(m + q);
          const WarpSample s = make_warp_sample(mm, (float)(x0 + i), (float)(y0 + r), depth, p.H, p.W, C);
          sh.tab[warp][lane] = s;
          unsigned sl = 0;
          if (s.p00 != kNoSample) {
            const unsigned cxl = (s.p00 / (unsigned)C) % (unsigned)p.W;
            const unsigned cxr = (s.p01 / (unsigned)C) % (unsigned)p.W;
            sl = (cxl % kBSlots) | ((cxr % kBSlots) << 8);
          }
          sh.slots[warp][lane] = sl;
        }
        __syncwarp();

        P4 open_top[KMAX], open_bot[KMAX];
        unsigned o_top[KMAX], o_bot[KMAX], s_top[KMAX], s_bot[KMAX];
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          o_top[j] = o_bot[j] = kNoCell;
          s_top[j] = s_bot[j] = 0;
          open_top[j] = open_bot[j] = p4zero();
        }
        const TG* g_row = g_d + (size_t)r * p.W * C;
        const TIn* ref_row = ref_blk + (size_t)r * p.W * C;
#pragma unroll 1
        for (int i = 0; i < npix; ++i) {
          const typename Raw<TG>::type graw = Raw<TG>::ld_stream(g_row + i * C);
          const typename Raw<TIn>::type rraw = Raw<TIn>::ld(ref_row + i * C);
          WarpSample smp[KMAX];
          P4 wv[KMAX];
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            wv[j] = p4zero();
            if (j >= k) continue;
            smp[j] = sh.tab[warp][i * k + j];
            if (smp[j].p00 == kNoSample) continue;
            P4 w1[1];
            gather_taps_p<TIn, 1, true>(nsrc[j], smp[j], 0, C, w1);
            wv[j] = w1[0];
          }
          const P4 ref = p4from(rraw);
          P4 mu = ref;
#pragma unroll
          for (int j = 0; j < KMAX; ++j)
            if (j < k) mu = p4add(mu, wv[j]);
          mu = p4scale(mu, inv_n2);
          const P4 gv = p4scale(p4from(graw), two_inv_n2);
          {   // reference gradient, accumulated over planes in tensor memory
            const uint32_t ta = tbase + 4u * (uint32_t)(r * kBRun + i);
            tmem_st4(ta, p4fma(gv, p4sub(ref, mu), tmem_ld4(ta)));
          }
#pragma unroll
          for (int j = 0; j < KMAX; ++j) {
            if (j >= k) continue;
            const WarpSample s = smp[j];
            if (s.p00 == kNoSample) {
              blk_flush<true>(on, sh, warp, lane, j, ndst[j], o_top[j], s_top[j], open_top[j]);
              blk_flush<false>(on, sh, warp, lane, j, ndst[j], o_bot[j], s_bot[j], open_bot[j]);
              continue;
            }
            const unsigned sl = sh.slots[warp][i * k + j];
            const P4 gw = p4mul(gv, p4sub(wv[j], mu));
            blk_side<true>(on, sh, warp, lane, j, ndst[j], gw, s.w00, s.w01, s.p00, s.p01, sl & 0xff,
                                 sl >> 8, o_top[j], s_top[j], open_top[j]);
            blk_side<false>(on, sh, warp, lane, j, ndst[j], gw, s.w10, s.w11, s.p10, s.p11, sl & 0xff,
                                  sl >> 8, o_bot[j], s_bot[j], open_bot[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          if (j >= k) continue;
          blk_flush<true>(on, sh, warp, lane, j, ndst[j], o_top[j], s_top[j], open_top[j]);
          blk_flush<false>(on, sh, warp, lane, j, ndst[j], o_bot[j], s_bot[j], open_bot[j]);
        }
      }
      // ---- end of the block for this plane: drain the row cache
      __syncwarp();
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j >= k) continue;
        for (int s = 0; s < kBSlots; ++s) {
          const unsigned t = sh.tag[warp][j][s];
          if (t != kNoCell && on) red_p4(at(ndst[j], t), sh.cache[warp][j][s][lane]);
        }
      }
      __syncwarp();
      if (lane < 2 * kBSlots) sh.tag[warp][lane / kBSlots][lane % kBSlots] = kNoCell;
      tmem_wait_st();
      __syncwarp();
    }
    // ---- reference gradients out of tensor memory
    float* dst = p.g_feat + (size_t)(v + p.ref_begin) * HW * C + blk_off;
    for (int r = 0; r < nrows; ++r)
      for (int i = 0; i < npix; ++i) {
        const P4 g = tmem_ld4(tbase + 4u * (uint32_t)(r * kBRun + i));
        if (on) red_p4(dst + ((size_t)r * p.W + i) * C, g);
      }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
                 :: "r"(sh.tmem_base), "n"(kBTmemCols) : "memory");
}

template <typename TIn, typename TG>
static int launch_bwd_blk_t(SweepParams& p, cudaStream_t st) {
  p.slices = (p.C + 127) / 128;
  const int n_xr = (p.W + kBRun - 1) / kBRun, n_rb = (p.H + kBRows - 1) / kBRows;
  const long long items = (long long)p.V * p.slices * n_rb * n_xr;
  const long long blocks = (items + kBWarps - 1) / kBWarps;
  if (blocks > 2147483647LL) return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd: grid too large");
  const bool full = p.C % 128 == 0;
  const size_t smem = sizeof(BlkShared);
#define MVSD_BLK(KM, FU)                                                                           \
  do {                                                                                             \
    auto kern = sweep_bwd_blk_kernel<TIn, TG, KM, FU>;                                             \
    static bool attr_set = false;                                                                  \
    if (!attr_set) {                                                                               \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
      attr_set = true;                                                                             \
    }                                                                                              \
    kern<<<(unsigned)blocks, kBThreads, smem, st>>>(p, (int)items, n_xr, n_rb);                    \
  } while (0)
  if (p.k == 1) { if (full) MVSD_BLK(1, true); else MVSD_BLK(1, false); }
  else { if (full) MVSD_BLK(2, true); else MVSD_BLK(2, false); }
#undef MVSD_BLK
  count_launch();
  return check_launch("plane_sweep_bwd(block)");
}

int launch_bwd_blk(SweepParams& p, int feat_dtype, int g_dtype, cudaStream_t st) {
  if (p.k >= 1 && p.k <= 2) {
    if (feat_dtype == MVSD_F32 && g_dtype == MVSD_F32) return launch_bwd_blk_t<float, float>(p, st);
    if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_F32)
      return launch_bwd_blk_t<__nv_bfloat16, float>(p, st);
    if (feat_dtype == MVSD_BF16 && g_dtype == MVSD_BF16)
      return launch_bwd_blk_t<__nv_bfloat16, __nv_bfloat16>(p, st);
  }
  return fail(MVSD_ERR_UNSUPPORTED, "plane_sweep_bwd(block): configuration not built");
}

}  // namespace mvsd
