// prep.cu -- per-step prologue kernel: image pyramid (F.resize_images, base_model.py:70-72) staged as
// NHWC4 so that every bilinear tap of the warp is one 16-byte load, the 3x4 projection tables
// (proj_tgt_to_src, transform.py:64-91) and inverse intrinsics (F.batch_inv, transform.py:105), and the
// reset of the fp64 reduction cells the fused loss kernel accumulates into.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kPrepThreads = 256;
constexpr int kBandMax = 8;       // full-resolution rows per pyramid CTA (p.band = 8 / p.split)

// Pyramid CTA = (image, band of kBand full-resolution rows).  It copies the band to scale 0 and produces
// every coarser-scale row whose top tap row v0 lies in the band, so the full-resolution planes are read
// from HBM once (the coarser scales re-read them through L1/L2 while they are hot) instead of once per
// scale.  One texel per thread and iteration, lanes along x: 4-byte planar loads and 16-byte texel stores
// are fully coalesced.
// Coordinates follow Chainer's resize_images: u = linspace(0, W-1, w_s) in float64 (x*step, last
// element pinned to W-1), u0 = clip(floor(u), 0, W-2), weights are float64 products cast to fp32,
// y = ((w1*a + w2*b) + w3*c) + w4*d in fp32.  Scale 0 is the identity and is copied.
// Source images carry the zero border of the padded layout (common.cuh): column w of every row and the
// two rows below the image.
template <int SPLIT>
__global__ void __launch_bounds__(kPrepThreads) sfm_prep_kernel(const __grid_constant__ SfmPrepParams p) {
  const int blk = blockIdx.x;
  cudaTriggerProgrammaticLaunchCompletion();      // the smoothness kernel may start once every CTA of this grid runs
  if (blk >= p.n_pyr_blocks) {
    // ---- tables + accumulator reset (a handful of trailing CTAs)
    const int t = (blk - p.n_pyr_blocks) * kPrepThreads + threadIdx.x;
    const int n_proj = p.build_tables ? p.B * p.S * p.ns : 0;
    const int n_kinv = p.build_tables ? p.B * p.ns : 0;
    if (t < n_proj) {
      const int s = t % p.ns;
      const int i = (t / p.ns) % p.S;
      const int b = t / (p.ns * p.S);
      float K[9], pose[6], P[12];
      for (int k = 0; k < 9; ++k) K[k] = p.intrinsics[((size_t)b * p.ns + s) * 9 + k];
      if (p.raw_pose_hw > 0) {
        // producer-side fusion: pose = 0.01 * mean_{h', w'}(poseout)   (pose_net.py:52-53; channel 6i+k of snippet b)
        for (int k = 0; k < 6; ++k)
          pose[k] = sfm_pose_component(p.poses + (((size_t)b * p.S + i) * 6 + k) * p.raw_pose_hw, p.raw_pose_hw);
        if (s == 0 && p.posevec_out)
          for (int k = 0; k < 6; ++k) p.posevec_out[((size_t)b * p.S + i) * 6 + k] = pose[k];
      } else {
        for (int k = 0; k < 6; ++k) pose[k] = p.poses[((size_t)b * p.S + i) * 6 + k];
      }
      sfm_build_proj(pose, K, P);
      for (int k = 0; k < 12; ++k) p.proj_out[(size_t)t * 12 + k] = P[k];
    } else if (t < n_proj + n_kinv) {
      const int j = t - n_proj;
      float K[9], inv[9];
      for (int k = 0; k < 9; ++k) K[k] = p.intrinsics[(size_t)j * 9 + k];
      sfm_inv3(K, inv);
      for (int k = 0; k < 9; ++k) p.kinv_out[(size_t)j * 9 + k] = inv[k];
    } else {
      const int j = t - n_proj - n_kinv;
      if (j < p.n_acc) p.acc[j] = 0.0;
      if (j == p.n_acc && p.counter) *p.counter = 0u;
    }
    return;
  }

  // ---- pyramid
  // p.split warps share one row (x interleaved in units of 32 texels): small batches are latency-bound here -- a warp
  // walks its scale-0 row and then its rows of the coarser scales one 32-texel iteration at a time -- so they get
  // shorter walks (and p.band = 8 / p.split rows per CTA) instead of idle SMs
  constexpr int kBand = kBandMax / SPLIT, split = SPLIT;
  const int n_bands = (p.H + kBand - 1) / kBand;
  int img = blk / n_bands;                  // [0, B): target b ; [B, B + B*S): source (b, i)
  const int band = blk - img * n_bands;
  if (p.interleave) {                       // snippet-major order: (target b, sources (b, 0..S-1)), b ascending
    const int b = img / (1 + p.S), j = img - b * (1 + p.S);
    img = (j == 0) ? b : p.B + b * p.S + (j - 1);
  }
  const int Y0 = band * kBand, Y1 = min(Y0 + kBand, p.H);
  const bool is_src = img >= p.B;
  const int H = p.H, W = p.W;
  const size_t plane = (size_t)H * W;
  const float* __restrict__ base = is_src ? p.src + (size_t)(img - p.B) * 3 * plane : p.tgt + (size_t)img * 3 * plane;
  const int warp = (threadIdx.x >> 5) / split, sub = (threadIdx.x >> 5) % split, lane = threadIdx.x & 31;
  constexpr int nrow = (kPrepThreads / 32) / split;       // rows walked concurrently by the CTA
  constexpr int x_step = 32 * split;
  const int x_first = lane + 32 * sub;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < p.ns; ++s) {
    const int h = H >> s, w = W >> s;
    const int pitch = is_src ? sfm_src_pitch(w) : w;
    float4* __restrict__ out = is_src ? p.src_pyr[s] + (size_t)(img - p.B) * sfm_src_rows(h) * pitch
                                      : p.tgt_pyr[s] + (size_t)img * h * w;
    if (s == 0) {
      for (int y = Y0 + warp; y < Y1; y += nrow) {
        const float* __restrict__ row = base + (size_t)y * W;
        float4* __restrict__ orow = out + (size_t)y * pitch;
        for (int x = x_first; x < W; x += x_step)
          orow[x] = make_float4(__ldg(row + x), __ldg(row + plane + x), __ldg(row + 2 * plane + x), 0.f);
        if (is_src && lane == 0 && sub == 0) orow[W] = zero4;
      }
    } else {
      const double stepx = (w > 1) ? __ddiv_rn((double)(W - 1), (double)(w - 1)) : 0.0;
      const double stepy = (h > 1) ? __ddiv_rn((double)(H - 1), (double)(h - 1)) : 0.0;
      // candidate rows: those whose v0 can fall in [Y0, Y1); the exact test is below
      const int y_lo = (stepy > 0.0) ? max(0, (int)((double)Y0 / stepy) - 1) : 0;
      const int y_hi = (stepy > 0.0) ? min(h, (int)((double)Y1 / stepy) + 2) : h;
      for (int y = y_lo + warp; y < y_hi; y += nrow) {
        const double v = (y == h - 1 && h > 1) ? (double)(H - 1) : __dmul_rn((double)y, stepy);
        const int v0 = min(max((int)floor(v), 0), H - 2);
        if (v0 < Y0 || v0 >= Y1) continue;          // another band owns this row (warp-uniform)
        const double va = __dsub_rn((double)(v0 + 1), v), vb = __dsub_rn(v, (double)v0);
        const float* __restrict__ r0 = base + (size_t)v0 * W;
        float4* __restrict__ orow = out + (size_t)y * pitch;
        for (int x = x_first; x < w; x += x_step) {
          const double u = (x == w - 1 && w > 1) ? (double)(W - 1) : __dmul_rn((double)x, stepx);
          const int u0 = min(max((int)floor(u), 0), W - 2);
          const double ua = __dsub_rn((double)(u0 + 1), u), ub = __dsub_rn(u, (double)u0);
          const float w1 = (float)__dmul_rn(ua, va), w2 = (float)__dmul_rn(ub, va);
          const float w3 = (float)__dmul_rn(ua, vb), w4 = (float)__dmul_rn(ub, vb);
          float r[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* pl = r0 + c * plane + u0;
            r[c] = sfm_blend(w1, w2, w3, w4, __ldg(pl), __ldg(pl + 1), __ldg(pl + W), __ldg(pl + W + 1));
          }
          orow[x] = make_float4(r[0], r[1], r[2], 0.f);
        }
        if (is_src && lane == 0 && sub == 0) orow[w] = zero4;
      }
    }
    // the two zero rows below a source image belong to the last band
    if (is_src && band == n_bands - 1)
      for (int k = threadIdx.x; k < 2 * pitch; k += kPrepThreads) out[(size_t)h * pitch + k] = zero4;
  }
}

__global__ void sfm_disp_activation_kernel(const float* __restrict__ x, float* __restrict__ disp, float* __restrict__ dact,
                                           long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float da;
    const float d = sfm_disp_act(__ldg(x + i), da);
    if (disp) disp[i] = d;
    if (dact) dact[i] = da;
  }
}

__global__ void sfm_pose_reduce_kernel(const float* __restrict__ x, float* __restrict__ out, int n_comp, int hw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_comp) out[t] = sfm_pose_component(x + (size_t)t * hw, hw);
}

__global__ void sfm_pyramid_export_kernel(const float4* __restrict__ pyr, float* __restrict__ out, long long n_img,
                                          int h, int w, int pitch, int rows) {
  const int hw = h * w;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_img * hw) return;
  const long long img = gid / hw;
  const int pix = (int)(gid - img * hw);
  const int y = pix / w, x = pix - y * w;
  const float4 v = pyr[img * ((long long)rows * pitch) + (long long)y * pitch + x];
  float* o = out + (size_t)img * 3 * hw + pix;
  o[0] = v.x;
  o[hw] = v.y;
  o[2 * (size_t)hw] = v.z;
}

}  // namespace

int sfm_launch_prep(const SfmPrepParams& p_in, cudaStream_t stream) {
  SfmPrepParams p = p_in;
  // split the rows over 2 or 4 warps while the grid stays within one wave of 256-thread CTAs (4 per SM)
  p.split = 1;
  {
    const char* e = getenv("SFM_PREP_SPLIT");     // development knob
    if (e && (atoi(e) == 1 || atoi(e) == 2 || atoi(e) == 4)) p.split = atoi(e);
    else
      while (p.split < 4 && (long long)p.B * (1 + p.S) * ((p.H * 2 * p.split + kBandMax - 1) / kBandMax) <= 148 * 4) p.split *= 2;
  }
  {
    const char* e = getenv("SFM_LIFO");           // development knob (see sfm_launch_fused)
    p.interleave = (e && atoi(e) > 0) ? 1 : 0;
  }
  p.band = kBandMax / p.split;
  const int kBand = p.band;
  p.n_pyr_blocks = p.do_pyramid ? p.B * (1 + p.S) * ((p.H + kBand - 1) / kBand) : 0;
  const long long n_tail = (p.build_tables ? (long long)p.B * p.S * p.ns + (long long)p.B * p.ns : 0) + p.n_acc + 1;
  const int tail_blocks = (int)((n_tail + kPrepThreads - 1) / kPrepThreads);
  if (p.split == 4) sfm_prep_kernel<4><<<p.n_pyr_blocks + tail_blocks, kPrepThreads, 0, stream>>>(p);
  else if (p.split == 2) sfm_prep_kernel<2><<<p.n_pyr_blocks + tail_blocks, kPrepThreads, 0, stream>>>(p);
  else sfm_prep_kernel<1><<<p.n_pyr_blocks + tail_blocks, kPrepThreads, 0, stream>>>(p);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sfm_launch_disp_activation(long long n, const float* x, float* disp, float* dact, cudaStream_t stream) {
  if (n <= 0) return 0;
  const long long blocks = (n + 255) / 256;
  sfm_disp_activation_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, stream>>>(x, disp, dact, n);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sfm_launch_pose_reduce(int B, int S, int hw, const float* x, float* poses_out, cudaStream_t stream) {
  const int n = B * S * 6;
  sfm_pose_reduce_kernel<<<(n + 127) / 128, 128, 0, stream>>>(x, poses_out, n, hw);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int sfm_launch_pyramid_export(const float4* pyr, float* out, long long n_img, int h, int w, int padded,
                              cudaStream_t stream) {
  const long long n = n_img * h * w;
  if (n == 0) return 0;
  sfm_pyramid_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pyr, out, n_img, h, w, padded ? sfm_src_pitch(w) : w,
                                                                           padded ? sfm_src_rows(h) : h);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}
