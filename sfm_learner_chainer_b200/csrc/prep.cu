// prep.cu -- per-step prologue kernel: image pyramid (F.resize_images, base_model.py:70-72) staged as
// NHWC4 so that every bilinear tap of the warp is one 16-byte load, the 3x4 projection tables
// (proj_tgt_to_src, transform.py:64-91) and inverse intrinsics (F.batch_inv, transform.py:105), and the
// reset of the fp64 reduction cells the fused loss kernel accumulates into.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kPrepThreads = 256;

// One thread = one output pixel of one image at one scale.
// Coordinates follow Chainer's resize_images: u = linspace(0, W-1, w_s) in float64 (x*step, last
// element pinned to W-1), u0 = clip(floor(u), 0, W-2), weights are float64 products cast to fp32,
// y = ((w1*a + w2*b) + w3*c) + w4*d in fp32.  Scale 0 is the identity and is copied.
__global__ void __launch_bounds__(kPrepThreads) sfm_prep_kernel(const __grid_constant__ SfmPrepParams p) {
  const int blk = blockIdx.x;
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
      for (int k = 0; k < 6; ++k) pose[k] = p.poses[((size_t)b * p.S + i) * 6 + k];
      sfm_build_proj(pose, K, P);
      for (int k = 0; k < 12; ++k) p.proj_out[(size_t)t * 12 + k] = P[k];
    } else if (t < n_proj + n_kinv) {
      const int j = t - n_proj;
      float K[9], inv[9];
      for (int k = 0; k < 9; ++k) K[k] = p.intrinsics[(size_t)j * 9 + k];
      sfm_inv3(K, inv);
      for (int k = 0; k < 9; ++k) p.kinv_out[(size_t)j * 9 + k] = inv[k];
    } else {
      long long j = (long long)t - n_proj - n_kinv;
      if (j < p.n_acc) { p.acc[j] = 0.0; return; }
      j -= p.n_acc;
      if (j == 0 && p.counter) *p.counter = 0u;
      if (!p.do_pyramid) return;
      // zero padding of the source pyramid: column w of every row, and the two rows below every image
      for (int s = 0; s < p.ns; ++s) {
        const int h = p.H >> s, w = p.W >> s, pitch = sfm_src_pitch(w);
        const long long per_img = (long long)h + 2ll * pitch;
        const long long n = (long long)p.B * p.S * per_img;
        if (j < n) {
          const long long img = j / per_img;
          const int k = (int)(j - img * per_img);
          const long long off = (k < h) ? (long long)k * pitch + w : (long long)h * pitch + (k - h);
          p.src_pyr[s][img * ((long long)sfm_src_rows(h) * pitch) + off] = make_float4(0.f, 0.f, 0.f, 0.f);
          return;
        }
        j -= n;
      }
    }
    return;
  }
  if (!p.do_pyramid) return;

  // ---- pyramid
  long long gid = (long long)blk * kPrepThreads + threadIdx.x;
  int s = 0;
#pragma unroll
  for (int k = 1; k < SFM_MAX_SCALES; ++k)
    if (k < p.ns && gid >= p.pix_begin[k]) s = k;
  gid -= p.pix_begin[s];
  const int h = p.H >> s, w = p.W >> s;
  const long long hw = (long long)h * w;
  const long long n_img = (long long)p.B * (1 + p.S);
  const size_t plane = (size_t)p.H * p.W;
  if (s == 0 && p.vec0) {
    // scale 0 is the identity: 4 pixels per thread, 3 x 16-byte planar loads -> 4 x 16-byte texel stores
    const long long q = hw / 4;
    if (gid >= n_img * q) return;
    const int img = (int)(gid / q);
    const int pix = (int)(gid - (long long)img * q) * 4;
    const float* base = (img < p.B) ? p.tgt + (size_t)img * 3 * plane : p.src + (size_t)(img - p.B) * 3 * plane;
    float4* out;
    if (img < p.B) {
      out = p.tgt_pyr[0] + (size_t)img * hw + pix;
    } else {
      const int yy = pix / w, xx = pix - yy * w;          // w % 4 == 0: the 4 pixels share a row
      out = p.src_pyr[0] + (size_t)(img - p.B) * sfm_src_rows(h) * sfm_src_pitch(w) + (size_t)yy * sfm_src_pitch(w) + xx;
    }
    const float4 r = __ldg(reinterpret_cast<const float4*>(base + pix));
    const float4 g = __ldg(reinterpret_cast<const float4*>(base + plane + pix));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(base + 2 * plane + pix));
    out[0] = make_float4(r.x, g.x, bb.x, 0.f);
    out[1] = make_float4(r.y, g.y, bb.y, 0.f);
    out[2] = make_float4(r.z, g.z, bb.z, 0.f);
    out[3] = make_float4(r.w, g.w, bb.w, 0.f);
    return;
  }
  if (gid >= n_img * hw) return;
  const int img = (int)(gid / hw);          // [0, B): target b ; [B, B + B*S): source (b, i)
  const int pix = (int)(gid - (long long)img * hw);
  const int y = pix / w, x = pix - y * w;
  const float* base;
  float4* out;
  if (img < p.B) {
    base = p.tgt + (size_t)img * 3 * plane;
    out = p.tgt_pyr[s] + (size_t)img * hw + pix;
  } else {
    base = p.src + (size_t)(img - p.B) * 3 * plane;
    out = p.src_pyr[s] + (size_t)(img - p.B) * sfm_src_rows(h) * sfm_src_pitch(w) + (size_t)y * sfm_src_pitch(w) + x;
  }
  float4 o;
  o.w = 0.f;
  if (s == 0) {
    const size_t a = (size_t)y * p.W + x;
    o.x = __ldg(base + a);
    o.y = __ldg(base + plane + a);
    o.z = __ldg(base + 2 * plane + a);
  } else {
    const double stepx = (w > 1) ? __ddiv_rn((double)(p.W - 1), (double)(w - 1)) : 0.0;
    const double stepy = (h > 1) ? __ddiv_rn((double)(p.H - 1), (double)(h - 1)) : 0.0;
    const double u = (x == w - 1 && w > 1) ? (double)(p.W - 1) : __dmul_rn((double)x, stepx);
    const double v = (y == h - 1 && h > 1) ? (double)(p.H - 1) : __dmul_rn((double)y, stepy);
    int u0 = (int)floor(u), v0 = (int)floor(v);
    u0 = min(max(u0, 0), p.W - 2);
    v0 = min(max(v0, 0), p.H - 2);
    const double ua = __dsub_rn((double)(u0 + 1), u), ub = __dsub_rn(u, (double)u0);
    const double va = __dsub_rn((double)(v0 + 1), v), vb = __dsub_rn(v, (double)v0);
    const float w1 = (float)__dmul_rn(ua, va), w2 = (float)__dmul_rn(ub, va);
    const float w3 = (float)__dmul_rn(ua, vb), w4 = (float)__dmul_rn(ub, vb);
    const size_t a00 = (size_t)v0 * p.W + u0;
    const size_t a10 = a00 + p.W;
    float r[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = base + c * plane;
      r[c] = sfm_blend(w1, w2, w3, w4, __ldg(pl + a00), __ldg(pl + a00 + 1), __ldg(pl + a10), __ldg(pl + a10 + 1));
    }
    o.x = r[0];
    o.y = r[1];
    o.z = r[2];
  }
  *out = o;
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
  // vector path for scale 0 needs 16-byte aligned planes and rows
  p.vec0 = (p.do_pyramid && p.W % 4 == 0 && ((uintptr_t)p.tgt & 15) == 0 && ((uintptr_t)p.src & 15) == 0) ? 1 : 0;
  long long total = 0;
  for (int s = 0; s < SFM_MAX_SCALES; ++s) {
    p.pix_begin[s] = total;
    if (s < p.ns && p.do_pyramid) {
      const long long n = (long long)p.B * (1 + p.S) * (p.H >> s) * (p.W >> s);
      total += (s == 0 && p.vec0) ? n / 4 : n;
    }
  }
  p.n_pyr_blocks = (int)((total + kPrepThreads - 1) / kPrepThreads);
  long long n_pad = 0;
  if (p.do_pyramid)
    for (int s = 0; s < p.ns; ++s) n_pad += (long long)p.B * p.S * ((p.H >> s) + 2ll * sfm_src_pitch(p.W >> s));
  const long long n_tail = (p.build_tables ? (long long)p.B * p.S * p.ns + (long long)p.B * p.ns : 0) + p.n_acc + 1 + n_pad;
  const int tail_blocks = (int)((n_tail + kPrepThreads - 1) / kPrepThreads);
  sfm_prep_kernel<<<p.n_pyr_blocks + tail_blocks, kPrepThreads, 0, stream>>>(p);
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
