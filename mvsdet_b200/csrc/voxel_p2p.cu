// Voxel-volume combine of a view-sharded scene over NVLink peer memory.
//
// After the local back-projection every rank holds partial per-voxel feature sums and
// valid counts for its reference views (mvsdet.py:1458-1460 per view); the reference then
// sums over views and divides by the count (mvsdet.py:511-515, :681-682).  Across ranks
// that is sum-all-reduce + normalise.  Instead of NCCL all-reduce + a normalise kernel
// (three passes over 26 MB and a latency-bound collective), ONE kernel does the
// reduce-scatter, the normalisation and the all-gather over peer pointers:
//   rank r owns the r-th 1/G of the flat [C*N] volume; for each owned float4 it loads the
//   partial from every peer (NVLink P2P loads, rank order -> identical bits on every
//   rank), divides by the reduced count and stores the result into EVERY peer's output
//   buffer (P2P stores).  Per rank: (G-1)/G * 26 MB in and out over NVLink, nothing
//   through HBM twice.
// The caller brackets the kernel with two cross-rank barriers (partials complete /
// results delivered); buffers come from torch's symmetric-memory allocator.
#include "common.cuh"

namespace mvsd {
namespace {

constexpr int kMaxPeers = 16;

struct P2PParams {
  const float* const* part;     // device array [world]: peers' partial buffers ([total] fp32 sums, then [N] int32 counts)
  float* const* out;            // device array [world]: peers' result buffers (same layout)
  int32_t* count_local;         // [N] reduced counts (also written to this rank's out tail)
  int world, rank, cfirst, C, N;
};

__device__ __forceinline__ float4 ld_peer4(const float* p) {       // never from a stale L1 line
  return __ldcv(reinterpret_cast<const float4*>(p));
}

// every rank reduces ALL counts (N * world 4-byte loads: 0.1 MB per peer)
__global__ void p2p_count_kernel(const P2PParams p) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= p.N) return;
  const size_t total = (size_t)p.C * p.N;
  int cnt = 0;
  for (int q = 0; q < p.world; ++q)
    cnt += __ldcv(reinterpret_cast<const int32_t*>(p.part[q] + total) + u);
  p.count_local[u] = cnt;
  reinterpret_cast<int32_t*>(p.out[p.rank] + total)[u] = cnt;
}

template <int WORLD>            // 0 = run-time world size
__global__ void __launch_bounds__(256) p2p_reduce_kernel(const P2PParams p) {
  const size_t total4 = ((size_t)p.C * p.N) >> 2;                // float4 elements (C*N % 4 == 0)
  const int world = WORLD ? WORLD : p.world;
  const size_t chunk = (total4 + world - 1) / world;
  const size_t begin = chunk * p.rank, end = min(total4, begin + chunk);
  const float* part[kMaxPeers];
  float* out[kMaxPeers];
#pragma unroll
  for (int q = 0; q < (WORLD ? WORLD : kMaxPeers); ++q) {
    if (q < world) {
      part[q] = p.part[q];
      out[q] = p.out[q];
    }
  }
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i << 2;
    float4 s = ld_peer4(part[0] + e);
#pragma unroll
    for (int q = 1; q < (WORLD ? WORLD : kMaxPeers); ++q) {
      if (q < world) {
        const float4 t = ld_peer4(part[q] + e);
        s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y);
        s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
      }
    }
    int c0, c1, c2, c3;
    if (p.cfirst) {                                              // [C][N]: four consecutive voxels
      const int u = (int)(e % (size_t)p.N);
      const int4 c = *reinterpret_cast<const int4*>(p.count_local + u);
      c0 = c.x; c1 = c.y; c2 = c.z; c3 = c.w;
    } else {                                                     // [N][C]: one voxel, four channels
      c0 = c1 = c2 = c3 = p.count_local[e / (size_t)p.C];
    }
    // sum / (count + 1e-8), zero where count == 0 (mvsdet.py:514-515)
    float4 m;
    m.x = c0 ? __fdiv_rn(s.x, __fadd_rn((float)c0, 1e-8f)) : 0.f;
    m.y = c1 ? __fdiv_rn(s.y, __fadd_rn((float)c1, 1e-8f)) : 0.f;
    m.z = c2 ? __fdiv_rn(s.z, __fadd_rn((float)c2, 1e-8f)) : 0.f;
    m.w = c3 ? __fdiv_rn(s.w, __fadd_rn((float)c3, 1e-8f)) : 0.f;
#pragma unroll
    for (int q = 0; q < (WORLD ? WORLD : kMaxPeers); ++q)
      if (q < world) *reinterpret_cast<float4*>(out[q] + e) = m;
  }
}

// Backward of a view-sharded scene: a rank's plane-sweep backward scatters into the gradient maps
// of its block views AND of its halo views (neighbours owned by other ranks).  The owner of a view
// adds the halo contributions of its peers to its own map: one pull kernel over NVLink peer
// pointers, one CTA column per (owned view, 16-byte chunk), sources in table order -> the same
// summation order on every run (deterministic).  pulls: CSR over the owned views --
// offs[d] .. offs[d+1] index (src_rank, src_local_view) pairs.
struct HaloParams {
  float* const* g;            // device array [world]: peers' gradient buffers [V_local][view_elems] fp32
  const int32_t* offs;        // [n_own + 1]
  const int32_t* src;         // [n_pull][2]
  int rank, n_own;
  size_t view_elems;          // H*W*C, multiple of 4
};

__global__ void __launch_bounds__(256) halo_reduce_kernel(const HaloParams p) {
  const int d = blockIdx.y;
  const int b = p.offs[d], e = p.offs[d + 1];
  if (b == e) return;
  float* mine = p.g[p.rank] + (size_t)d * p.view_elems;
  const size_t n4 = p.view_elems >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = *reinterpret_cast<const float4*>(mine + (i << 2));
    for (int q = b; q < e; ++q) {
      const float* peer = p.g[p.src[2 * q]] + (size_t)p.src[2 * q + 1] * p.view_elems;
      const float4 t = ld_peer4(peer + (i << 2));
      s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y);
      s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
    }
    *reinterpret_cast<float4*>(mine + (i << 2)) = s;
  }
}

}  // namespace
}  // namespace mvsd

using namespace mvsd;

extern "C" int mvsd_halo_reduce_p2p(void* const* g_ptrs, const int32_t* pull_offsets,
                                    const int32_t* pull_sources, int world, int rank, int n_own,
                                    int64_t view_elems, void* stream) {
  if (!g_ptrs || !pull_offsets || !pull_sources)
    return fail(MVSD_ERR_INVALID_ARG, "halo_reduce_p2p: null pointer");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(MVSD_ERR_INVALID_ARG, "halo_reduce_p2p: bad world/rank (%d/%d)", rank, world);
  if (n_own <= 0 || n_own > 65535 || view_elems <= 0 || (view_elems & 3))
    return fail(MVSD_ERR_INVALID_ARG, "halo_reduce_p2p: bad view count / size");
  HaloParams p;
  p.g = reinterpret_cast<float* const*>(g_ptrs);
  p.offs = pull_offsets; p.src = pull_sources; p.rank = rank; p.n_own = n_own;
  p.view_elems = (size_t)view_elems;
  const size_t want = ((size_t)view_elems / 4 + 255) / 256;
  dim3 grid((unsigned)(want < 64 ? want : 64), (unsigned)n_own);
  halo_reduce_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  return check_launch("halo_reduce_p2p");
}

extern "C" int mvsd_voxel_reduce_p2p(const void* const* part_ptrs, void* const* out_ptrs,
                                     int32_t* count_local, int world, int rank, int layout, int C,
                                     int N, void* stream) {
  if (!part_ptrs || !out_ptrs || !count_local)
    return fail(MVSD_ERR_INVALID_ARG, "voxel_reduce_p2p: null pointer");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(MVSD_ERR_INVALID_ARG, "voxel_reduce_p2p: bad world/rank (%d/%d, at most %d peers)", rank,
                world, kMaxPeers);
  if (C <= 0 || N <= 0 || (C % 4) || (N % 4))
    return fail(MVSD_ERR_UNSUPPORTED, "voxel_reduce_p2p: C and N must be positive multiples of 4");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  P2PParams p;
  p.part = reinterpret_cast<const float* const*>(part_ptrs);
  p.out = reinterpret_cast<float* const*>(out_ptrs);
  p.count_local = count_local;
  p.world = world; p.rank = rank; p.cfirst = layout == MVSD_CHANNELS_FIRST ? 1 : 0; p.C = C; p.N = N;
  p2p_count_kernel<<<(N + 255) / 256, 256, 0, st>>>(p);
  count_launch();
  if (int e = check_launch("voxel_reduce_p2p(count)")) return e;
  // 148 SMs x 2 CTAs of 256 threads, each thread ~4 float4 per peer in flight: enough peer loads to
  // cover the NVLink round trip without taking every SM slot from a sweep that runs concurrently
  const size_t total4 = ((size_t)C * N) >> 2;
  const size_t chunk = (total4 + world - 1) / world;
  const size_t want = (chunk + 255) / 256;
  const unsigned blocks = (unsigned)(want < 148 * 2 ? (want ? want : 1) : 148 * 2);
  switch (world) {
    case 2: p2p_reduce_kernel<2><<<blocks, 256, 0, st>>>(p); break;
    case 4: p2p_reduce_kernel<4><<<blocks, 256, 0, st>>>(p); break;
    case 8: p2p_reduce_kernel<8><<<blocks, 256, 0, st>>>(p); break;
    default: p2p_reduce_kernel<0><<<blocks, 256, 0, st>>>(p); break;
  }
  count_launch();
  return check_launch("voxel_reduce_p2p");
}
