// Per-scene camera geometry in ONE launch.
//
// The reference rebuilds, for every scene and on the host / with ~35 tiny ATen calls
// (projects/NeRF-Det/nerfdet/mvsdet.py:407-434, :448-450, :1124-1156):
//   src_c2w = inverse(src_w2c); neighbour ids = k nearest camera centres   (:43-104, :432-434)
//   ref_proj = K_feat @ w2c;  neighbour projections gathered               (:249-264)
//   proj = nei_proj @ inverse(ref_proj) inside homo_warping               (module.py:116-118)
//   projection = K_feat[:3,:3] @ w2c[:3] per view                          (:1124-1156)
// Here the host keeps only what must be LAPACK's bits -- ref_proj = K_feat @ w2c and its
// inverse: the variance volume is sensitive to the rounding of that fp32 inverse (DESIGN.md
// section 6a) -- and uploads {w2c, K_feat, ref_proj, inverse(ref_proj)} in one pinned copy.
// This kernel produces the whole parameter block the sweep and back-projection kernels read:
//   nbr_ids    [n_ref,k]    int32   nearest first, self excluded
//   hom        [n_ref,k,12] fp32    rows of rot (9) then trans (3) of P_nbr @ inverse(P_ref)
//   projection [n_ref,3,4]  fp32
// Rounding follows ATen's CPU kernels, which is what the oracle (the reference on CPU) runs:
//   * batched 4x4 matmul (bmm of small matrices): out = ((a0*b0 + a1*b1) + a2*b2) + a3*b3 with
//     every product and sum rounded, NO fma (checked bit for bit on the build host);
//   * 2-D mm of a 3x3 by a 3x4: an FMA chain in k order (tests/test_geometry_cpu.py).
// Camera centres come from a closed-form inverse of w2c evaluated in fp64 and rounded to fp32
// (the reference's are LAPACK's fp32 inverse: equal to ~1 ulp); only their ORDERING matters.
// One CTA; a scene has at most a few hundred views.
#include "common.cuh"

namespace mvsd {

constexpr int kSetupMaxViews = 1024;
constexpr int kSetupThreads = 256;

struct SetupParams {
  const float* w2c;        // [V,16]
  const float* k_feat;     // [1 or V][16]
  const float* ref_proj;   // [V,16]
  const float* inv_ref;    // [V,16]
  int32_t* nbr;            // [n_ref,k]
  float* hom;              // [n_ref,k,12]
  float* projection;       // [n_ref,3,4]
  int V, k, per_view_k, ref_begin, n_ref;
};

// translation column of inverse(M) for a general 4x4 (cofactors, fp64)
__device__ __forceinline__ void inverse_translation(const float* m, double (&t)[3]) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (double)m[i];
  // Laplace expansion by 2x2 minors of rows {0,1} (s*) and rows {2,3} (c*)
  const double s0 = a[0] * a[5] - a[1] * a[4], s1 = a[0] * a[6] - a[2] * a[4];
  const double s2 = a[0] * a[7] - a[3] * a[4], s3 = a[1] * a[6] - a[2] * a[5];
  const double s4 = a[1] * a[7] - a[3] * a[5], s5 = a[2] * a[7] - a[3] * a[6];
  const double c5 = a[10] * a[15] - a[11] * a[14], c4 = a[9] * a[15] - a[11] * a[13];
  const double c3 = a[9] * a[14] - a[10] * a[13], c2 = a[8] * a[15] - a[11] * a[12];
  const double c1 = a[8] * a[14] - a[10] * a[12], c0 = a[8] * a[13] - a[9] * a[12];
  const double det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
  const double inv = 1.0 / det;
  t[0] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * inv;   // inverse[0][3]
  t[1] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * inv;    // inverse[1][3]
  t[2] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * inv;    // inverse[2][3]
}

__global__ void __launch_bounds__(kSetupThreads) scene_setup_kernel(const SetupParams p) {
  __shared__ float s_loc[kSetupMaxViews][3];
  const int tid = threadIdx.x;
  for (int v = tid; v < p.V; v += kSetupThreads) {
    double t[3];
    inverse_translation(p.w2c + (size_t)v * 16, t);
    s_loc[v][0] = (float)t[0];
    s_loc[v][1] = (float)t[1];
    s_loc[v][2] = (float)t[2];
  }
  __syncthreads();
  // ---- neighbour ids: k smallest squared centre distances, self masked (mvsdet.py:43-104)
  for (int r = tid; r < p.n_ref; r += kSetupThreads) {
    const int v = p.ref_begin + r;
    const double x = s_loc[v][0], y = s_loc[v][1], z = s_loc[v][2];
    double best[MVSD_MAX_K];
    int bid[MVSD_MAX_K];
#pragma unroll
    for (int j = 0; j < MVSD_MAX_K; ++j) { best[j] = 1e300; bid[j] = -1; }
    for (int n = 0; n < p.V; ++n) {
      if (n == v) continue;
      const double dx = (double)s_loc[n][0] - x, dy = (double)s_loc[n][1] - y, dz = (double)s_loc[n][2] - z;
      double d2 = dx * dx + dy * dy + dz * dz;
      int id = n;
#pragma unroll
      for (int j = 0; j < MVSD_MAX_K; ++j) {           // insertion, strict '<': ties keep the lower id first
        if (j < p.k && d2 < best[j]) {
          const double td = best[j]; const int ti = bid[j];
          best[j] = d2; bid[j] = id;
          d2 = td; id = ti;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < MVSD_MAX_K; ++j)
      if (j < p.k) p.nbr[(size_t)r * p.k + j] = bid[j];
  }
  // ---- projection = K_feat[:3,:3] @ w2c[:3]: FMA chain in k order (ATen 2-D mm)
  for (int e = tid; e < p.n_ref * 12; e += kSetupThreads) {
    const int r = e / 12, q = e - r * 12, i = q >> 2, c = q & 3;
    const int v = p.ref_begin + r;
    const float* K = p.k_feat + (p.per_view_k ? (size_t)v * 16 : 0);
    const float* E = p.w2c + (size_t)v * 16;
    float acc = __fmul_rn(K[i * 4 + 0], E[0 * 4 + c]);
    acc = fmaf(K[i * 4 + 1], E[1 * 4 + c], acc);
    acc = fmaf(K[i * 4 + 2], E[2 * 4 + c], acc);
    p.projection[e] = acc;
  }
  __syncthreads();          // nbr ids written above are read below (same CTA, global memory)
  // ---- hom = rows 0..2 of ref_proj[nbr] @ inv_ref[v]: un-fused chain (ATen batched matmul)
  const int per_ref = p.k * 12;
  for (int e = tid; e < p.n_ref * per_ref; e += kSetupThreads) {
    const int r = e / per_ref, q = e - r * per_ref, j = q / 12, o = q - j * 12;
    const int v = p.ref_begin + r;
    const int n = p.nbr[(size_t)r * p.k + j];
    const int i = o < 9 ? o / 3 : o - 9;              // row of M
    const int c = o < 9 ? o - 3 * (o / 3) : 3;        // column of M
    const float* A = p.ref_proj + (size_t)n * 16;
    const float* B = p.inv_ref + (size_t)v * 16;
    float acc = __fmul_rn(A[i * 4 + 0], B[0 * 4 + c]);
    acc = __fadd_rn(acc, __fmul_rn(A[i * 4 + 1], B[1 * 4 + c]));
    acc = __fadd_rn(acc, __fmul_rn(A[i * 4 + 2], B[2 * 4 + c]));
    acc = __fadd_rn(acc, __fmul_rn(A[i * 4 + 3], B[3 * 4 + c]));
    p.hom[e] = acc;
  }
}

}  // namespace mvsd

using namespace mvsd;

extern "C" int mvsd_scene_setup(const float* w2c, const float* k_feat, int per_view_k,
                                const float* ref_proj, const float* inv_ref, int32_t* nbr_ids,
                                float* hom, float* projection, int V, int k, int ref_begin,
                                int n_ref, void* stream) {
  if (V <= 0 || k < 0 || n_ref <= 0 || ref_begin < 0 || ref_begin + n_ref > V)
    return fail(MVSD_ERR_INVALID_ARG, "scene_setup: bad view counts (V=%d, k=%d, ref views [%d, %d))", V, k,
                ref_begin, ref_begin + n_ref);
  if (V > kSetupMaxViews) return fail(MVSD_ERR_UNSUPPORTED, "scene_setup: V=%d > %d", V, kSetupMaxViews);
  if (k > MVSD_MAX_K || k > V - 1)
    return fail(MVSD_ERR_UNSUPPORTED, "scene_setup: k=%d must be <= min(%d, V-1)", k, MVSD_MAX_K);
  if (!w2c || !k_feat || !projection || (k > 0 && (!ref_proj || !inv_ref || !nbr_ids || !hom)))
    return fail(MVSD_ERR_INVALID_ARG, "scene_setup: null pointer");
  SetupParams p{};
  p.w2c = w2c; p.k_feat = k_feat; p.ref_proj = ref_proj; p.inv_ref = inv_ref;
  p.nbr = nbr_ids; p.hom = hom; p.projection = projection;
  p.V = V; p.k = k; p.per_view_k = per_view_k ? 1 : 0; p.ref_begin = ref_begin; p.n_ref = n_ref;
  scene_setup_kernel<<<1, kSetupThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  return check_launch("scene_setup");
}
