// smooth.cu -- second-order disparity smoothness (compute_smooth_loss, base_model.py:169-185), forward
// and backward, for all snippets and scales in one launch.
//
//   dx = D[:, 1:] - D[:, :-1] ; dy = D[1:, :] - D[:-1, :]
//   loss_s = smooth_reg/2^s * ( mean|dx2| + mean|dxdy| + mean|dydx| + mean|dy2| )        (:181-185, :76)
//
// The term depends on the disparity alone, so it runs as its own small stencil kernel ahead of the
// fused photometric kernel: it writes gdisp completely (every pixel, zero where no term touches it) and
// the fused kernel then accumulates the photometric gradient on top.
//
// Marching formulation (same shape as the SSIM kernel): a warp owns a strip of 28 interior columns (+2
// halo columns per side) x hseg rows and streams the disparity rows y0-2 .. y1+1 through registers, lane =
// column.  Horizontal neighbours come from warp shuffles (6 per row), vertical ones from short register
// rings.  With S2x, S2y, M the signed weights of the second differences (0 where the difference does not
// exist) the gradient is the gather
//   G(y,x) = [S2x(y,x-2) - 2 S2x(y,x-1) + S2x(y,x)] + [S2y(y-2,x) - 2 S2y(y-1,x) + S2y(y,x)]
//          + [M(y,x) - M(y,x-1)] - [M(y-1,x) - M(y-1,x-1)]
// so there is no scatter and no atomic on gdisp.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int SM_IW = 28;         // interior columns per strip

__device__ __forceinline__ float sgnc(float v, float c) {   // sign(v) * c (c > 0), 0 when v == 0
  const float t = __int_as_float((__float_as_int(v) & 0x80000000) | __float_as_int(c));
  return (v == 0.f) ? 0.f : t;
}

template <bool GRAD>
__global__ void __launch_bounds__(32) sfm_smooth_kernel(const __grid_constant__ SfmFusedParams p) {
  const int lane = threadIdx.x;
  // No early cudaTriggerProgrammaticLaunchCompletion() here (the implicit trigger at CTA exit is used): the
  // dependent fused kernel has at most 12-20 single-warp CTAs per SM, and when its CTAs are dispatched while this
  // grid and the pyramid grid still occupy part of the chip they pile up on the SMs that happen to be free.  For
  // grids below one wave that imbalance sets the kernel time (measured at cfg2: step 58.6 us with the early
  // trigger vs 49.1 us without, same kernels).
#ifndef SFM_SMOOTH_TRIGGER
#define SFM_SMOOTH_TRIGGER 0
#endif
#if SFM_SMOOTH_TRIGGER
  cudaTriggerProgrammaticLaunchCompletion();
#endif
  // ---- task decode (uniform): strips x row segments of every (snippet, scale)
  int t = blockIdx.x, s = 0;
#pragma unroll
  for (int q = 1; q < SFM_MAX_SCALES; ++q)
    if (q < p.ns && t >= p.tile_begin[q]) s = q;
  t -= p.tile_begin[s];
  const int h = p.h[s], w = p.w[s];
  const int seg = t % p.tiles_y[s];
  t /= p.tiles_y[s];
  const int strip = t % p.tiles_x[s];
  const int b = t / p.tiles_x[s];
  const int y0 = seg * p.sm_hseg, y1 = min(y0 + p.sm_hseg, h);
  const int xx = strip * SM_IW - 2 + lane;
  const bool col_in = (xx >= 0) && (xx < w);
  const bool col_own = (lane >= 2) && (lane < 2 + SM_IW) && (xx < w);
  const float* __restrict__ D = p.disp[s] + (size_t)b * h * w;
  float* __restrict__ G = GRAD ? p.gdisp[s] + (size_t)b * h * w : nullptr;
  const float k_dx2 = p.sm_dx2[s], k_mix = p.sm_mix[s], k_dy2 = p.sm_dy2[s];
  const float gyv = (GRAD && p.gy) ? __ldg(p.gy) : 1.f;
  const bool x_dx2 = col_in && (xx <= w - 3);          // dx2(., xx) exists
  const bool x_mix = col_in && (xx <= w - 2);          // cell (., xx) exists
  float loss = 0.f;
  // rings (row index relative to the row r being loaded)
  float d1 = 0.f;                   // D[r-1]
  float ex1 = 0.f;                  // ex[r-1]
  float ey2 = 0.f;                  // ey[r-2] = D[r-1] - D[r-2]
  float sy3 = 0.f, sy4 = 0.f;       // S2y[r-3], S2y[r-4]
  float n2 = 0.f, n3 = 0.f;         // N[r-2], N[r-3],  N[y] = M(y,x) - M(y,x-1)
  float gx1 = 0.f, gx2 = 0.f;       // Gx[r-1], Gx[r-2]
  // producer-side fusion (SfmDesc.raw_disp_scales): D holds the pre-activation map, the disparity is formed on
  // load and the gradient is written w.r.t. the raw map (factor ring f0..f2 = d disp / d x of rows r..r-2)
  const bool raw = (p.raw_disp_mask >> s) & 1u;
  float f_next = 1.f, f1 = 1.f, f2 = 1.f;
  auto load = [&](int r, float& f) {
    f = 1.f;
    if (!(col_in && r >= 0 && r < h)) return 0.f;
    const float v = __ldg(D + (size_t)r * w + xx);
    return raw ? sfm_disp_act(v, f) : v;
  };
  float d_next = load(y0 - 2, f_next);
#pragma unroll 1
  for (int r = y0 - 2; r < y1 + 2; ++r) {
    const float d0 = d_next, f0 = f_next;
    d_next = load(r + 1, f_next);
    const bool r_in = (r >= 0) && (r < h);
    // ---- horizontal terms of row r
    const float ex0 = __fsub_rn(__shfl_down_sync(0xffffffffu, d0, 1), d0);                 // D[r][x+1] - D[r][x]
    const float dx2 = __fsub_rn(__shfl_down_sync(0xffffffffu, ex0, 1), ex0);
    const bool vx = r_in && x_dx2;
    const float sx = vx ? sgnc(dx2, k_dx2) : 0.f;
    if (vx && col_own && r >= y0 && r < y1) loss += fabsf(dx2) * k_dx2;
    const float sxm1 = __shfl_up_sync(0xffffffffu, sx, 1), sxm2 = __shfl_up_sync(0xffffffffu, sx, 2);
    const float gx0 = (sxm2 - 2.f * sxm1) + sx;
    // ---- vertical terms: dy2 at row r-2
    const float ey1 = __fsub_rn(d0, d1);                                                    // D[r] - D[r-1]
    const float dy2 = __fsub_rn(ey1, ey2);
    const int ry = r - 2;
    const bool vy = col_in && (ry >= 0) && (ry <= h - 3);
    const float sy2 = vy ? sgnc(dy2, k_dy2) : 0.f;
    if (vy && col_own && ry >= y0 && ry < y1) loss += fabsf(dy2) * k_dy2;
    // ---- mixed terms of cell (r-1, x): dxdy = ex[r] - ex[r-1] ; dydx = ey[r-1][x+1] - ey[r-1][x]
    const float a = __fsub_rn(ex0, ex1);
    const float bq = __fsub_rn(__shfl_down_sync(0xffffffffu, ey1, 1), ey1);
    const int rm = r - 1;
    const bool vm = x_mix && (rm >= 0) && (rm <= h - 2);
    const float m1 = vm ? (sgnc(a, k_mix) + sgnc(bq, k_mix)) : 0.f;
    if (vm && col_own && rm >= y0 && rm < y1) loss += (fabsf(a) + fabsf(bq)) * k_mix;
    const float n1 = m1 - __shfl_up_sync(0xffffffffu, m1, 1);
    // ---- gradient of pixel (r-2, x)
    if (GRAD) {
      const float g = (gx2 + ((sy4 - 2.f * sy3) + sy2)) + (n2 - n3);
      if (col_own && ry >= y0 && ry < y1) G[(size_t)ry * w + xx] = raw ? (gyv * g) * f2 : gyv * g;
    }
    f2 = f1; f1 = f0;
    d1 = d0; ex1 = ex0; ey2 = ey1;
    sy4 = sy3; sy3 = sy2;
    n3 = n2; n2 = n1;
    gx2 = gx1; gx1 = gx0;
  }
  loss = sfm_warp_sum(loss);
  // everything above depends only on the caller's disparity; the loss cell is reset by the prep kernel
  cudaGridDependencySynchronize();
  if (lane == 0 && loss != 0.f) atomicAdd(p.acc + 1, (double)loss);
}

// Edge-aware first-order smoothness, compute_disp_smooth (base_model.py:144-155; the alternative the reference
// keeps commented out at its call site, :78-80; SfmDesc flag SFM_FLAG_EDGE_AWARE_SMOOTH):
//   loss_s = smooth_reg/2^s * ( mean(|d_dx| * exp(-|mean_c i_dx|)) + mean(|d_dy| * exp(-|mean_c i_dy|)) )
// with d = disparity, i = target image of the scale (the pyramid level, no gradient).  One thread per pixel in
// gather form: its right / down differences give the loss terms it owns, and together with the left / up ones
// the gradient, so there is no scatter and no atomic on gdisp.  Needs the pyramid, so unlike the second-order
// kernel above it starts after the prep kernel.
template <bool GRAD>
__global__ void __launch_bounds__(256) sfm_edge_smooth_kernel(const __grid_constant__ SfmFusedParams p) {
  const int s = blockIdx.y % p.ns, b = blockIdx.y / p.ns;
  const int h = p.h[s], w = p.w[s], plane = h * w;
  const bool raw = (p.raw_disp_mask >> s) & 1u;
  const float* __restrict__ D = p.disp[s] + (size_t)b * plane;
  const float* __restrict__ T = p.tgt_pl[s] + (size_t)b * 3 * plane;      // planar (3, h, w)
  float* __restrict__ G = GRAD ? p.gdisp[s] + (size_t)b * plane : nullptr;
  const float kx = p.sm_ex[s], ky = p.sm_ey[s];
  const float gyv = (GRAD && p.gy) ? __ldg(p.gy) : 1.f;
  cudaGridDependencySynchronize();          // pyramid and loss cells (prep kernel)
  auto disp_at = [&](int i, float& f) {
    const float v = __ldg(D + i);
    f = 1.f;
    return raw ? sfm_disp_act(v, f) : v;
  };
  // exp(-|mean_c (b - a)|): F.mean over the three channels = ((d0 + d1) + d2) / 3, F.absolute, F.exp
  struct Px { float x, y, z; };
  auto tex = [&](int i) { Px v; v.x = __ldg(T + i); v.y = __ldg(T + plane + i); v.z = __ldg(T + 2 * plane + i); return v; };
  auto edge = [&](const Px& a, const Px& c) {
    const float m = __fdiv_rn(__fadd_rn(__fadd_rn(__fsub_rn(c.x, a.x), __fsub_rn(c.y, a.y)), __fsub_rn(c.z, a.z)), 3.f);
    return expf(-fabsf(m));
  };
  float loss = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
    const int y = i / w, x = i - y * w;
    float f0, fd;
    const float d0 = disp_at(i, f0);
    const Px t0 = tex(i);
    float g = 0.f;
    if (x + 1 < w) {
      const float dd = __fsub_rn(disp_at(i + 1, fd), d0), e = edge(t0, tex(i + 1));
      loss += fabsf(dd) * e * kx;
      g -= sgnc(dd, kx) * e;
    }
    if (y + 1 < h) {
      const float dd = __fsub_rn(disp_at(i + w, fd), d0), e = edge(t0, tex(i + w));
      loss += fabsf(dd) * e * ky;
      g -= sgnc(dd, ky) * e;
    }
    if (GRAD) {
      if (x > 0) g += sgnc(__fsub_rn(d0, disp_at(i - 1, fd)), kx) * edge(tex(i - 1), t0);
      if (y > 0) g += sgnc(__fsub_rn(d0, disp_at(i - w, fd)), ky) * edge(tex(i - w), t0);
      G[i] = raw ? (gyv * g) * f0 : gyv * g;
    }
  }
  __shared__ float part[8];
  loss = sfm_warp_sum(loss);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k];
    if (t != 0.f) atomicAdd(p.acc + 1, (double)t);
  }
}

}  // namespace

int sfm_launch_edge_smooth(SfmFusedParams& p, int grad, cudaStream_t stream) {
  const int per = (p.h[0] * p.w[0] + 255) / 256;
  dim3 grid((unsigned)(per > 32 ? 32 : per), (unsigned)(p.B * p.ns));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = sfm_pdl_enabled() ? attr : nullptr;
  cfg.numAttrs = sfm_pdl_enabled() ? 1 : 0;
  if (grad) SFM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sfm_edge_smooth_kernel<true>, p));
  else SFM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sfm_edge_smooth_kernel<false>, p));
  return 0;
}

// Fills the strip decomposition (tiles_x = strips, tiles_y = row segments, tile_begin) of `p` and launches.
// With grad != 0 every gdisp[s] is fully written.
int sfm_launch_smooth(SfmFusedParams& p, int grad, cudaStream_t stream) {
  if (p.edge_smooth) return sfm_launch_edge_smooth(p, grad, stream);
  // segment height: enough warps to cover the chip a few times, few enough that the 4 extra rows stay cheap
  long long strips = 0;
  for (int s = 0; s < p.ns; ++s) strips += (long long)p.B * ((p.w[s] + SM_IW - 1) / SM_IW);
  int hseg = 64;
  while (hseg > 8) {
    long long n = 0;
    for (int s = 0; s < p.ns; ++s) n += (long long)p.B * ((p.w[s] + SM_IW - 1) / SM_IW) * ((p.h[s] + hseg - 1) / hseg);
    if (n >= 148 * 4 * 8) break;
    hseg >>= 1;
  }
  p.sm_hseg = hseg;
  int total = 0;
  for (int s = 0; s < SFM_MAX_SCALES; ++s) {
    p.tile_begin[s] = total;
    if (s < p.ns) {
      p.tiles_x[s] = (p.w[s] + SM_IW - 1) / SM_IW;
      p.tiles_y[s] = (p.h[s] + hseg - 1) / hseg;
      total += p.B * p.tiles_x[s] * p.tiles_y[s];
    } else {
      p.tiles_x[s] = p.tiles_y[s] = 1;
    }
  }
  p.tile_begin[SFM_MAX_SCALES] = total;
  if (grad) SFM_CUDA_CHECK(sfm_launch_kernel(sfm_smooth_kernel<true>, total, 32, stream, true, p));
  else SFM_CUDA_CHECK(sfm_launch_kernel(sfm_smooth_kernel<false>, total, 32, stream, true, p));
  return 0;
}
