// smooth.cu -- second-order disparity smoothness (compute_smooth_loss, base_model.py:169-185), forward
// and backward, for all snippets and scales in one launch.
//
//   dx = D[:, 1:] - D[:, :-1] ; dy = D[1:, :] - D[:-1, :]
//   loss_s = smooth_reg/2^s * ( mean|dx2| + mean|dxdy| + mean|dydx| + mean|dy2| )        (:181-185, :76)
//
// The term depends on the disparity alone, so it runs as its own small stencil kernel ahead of the
// fused photometric kernel: it writes gdisp completely (every pixel, zero where no term touches it) and
// the fused kernel then accumulates the photometric gradient on top.  A CTA owns a 32x32 tile staged
// with a 2-pixel halo in shared memory; the gradient is evaluated in gather form (each pixel sums the
// signs of the second differences it takes part in), so there is no scatter and no atomic on gdisp.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TS = 32;            // tile edge
constexpr int HALO = 2;
constexpr int TP = TS + 2 * HALO; // padded tile edge

__device__ __forceinline__ float sgnc(float v, float c) {   // sign(v) * c (c > 0), 0 when v == 0
  const float t = __int_as_float((__float_as_int(v) & 0x80000000) | __float_as_int(c));
  return (v == 0.f) ? 0.f : t;
}

template <bool GRAD>
__global__ void __launch_bounds__(256) sfm_smooth_kernel(const __grid_constant__ SfmFusedParams p) {
  __shared__ float sd[TP][TP + 1];
  __shared__ float s_red[8];
  // ---- tile decode (uniform)
  int t = blockIdx.x, s = 0;
#pragma unroll
  for (int q = 1; q < SFM_MAX_SCALES; ++q)
    if (q < p.ns && t >= p.tile_begin[q]) s = q;
  t -= p.tile_begin[s];
  const int h = p.h[s], w = p.w[s];
  const int tx = t % p.tiles_x[s];
  t /= p.tiles_x[s];
  const int ty = t % p.tiles_y[s];
  const int b = t / p.tiles_y[s];
  const int x0 = tx * TS, y0 = ty * TS;
  const float* __restrict__ D = p.disp[s] + (size_t)b * h * w;
  for (int ly = threadIdx.x >> 5; ly < TP; ly += 8) {          // a warp per tile row: no index division
    const int yy = y0 - HALO + ly;
    const bool yin = (unsigned)yy < (unsigned)h;
    const float* __restrict__ row = D + (size_t)(yin ? yy : 0) * w;
    for (int lx = threadIdx.x & 31; lx < TP; lx += 32) {
      const int xx = x0 - HALO + lx;
      sd[ly][lx] = (yin && (unsigned)xx < (unsigned)w) ? __ldg(row + xx) : 0.f;
    }
  }
  __syncthreads();
  const float k_dx2 = p.sm_dx2[s], k_mix = p.sm_mix[s], k_dy2 = p.sm_dy2[s];
  const float gyv = (GRAD && p.gy) ? __ldg(p.gy) : 1.f;
  const int lx = (threadIdx.x & 31) + HALO;
  const int x = x0 + (threadIdx.x & 31);
  float loss = 0.f;
#pragma unroll
  for (int rr = 0; rr < TS / 8; ++rr) {
    const int ly = (threadIdx.x >> 5) + rr * 8 + HALO;
    const int y = y0 + ly - HALO;
    if (x < w && y < h) {
      const float c = sd[ly][lx];
      const float xm2 = sd[ly][lx - 2], xm1 = sd[ly][lx - 1], xp1 = sd[ly][lx + 1], xp2 = sd[ly][lx + 2];
      const float ym2 = sd[ly - 2][lx], ym1 = sd[ly - 1][lx], yp1 = sd[ly + 1][lx], yp2 = sd[ly + 2][lx];
      const float mm = sd[ly - 1][lx - 1], mp = sd[ly - 1][lx + 1], pm = sd[ly + 1][lx - 1], pp = sd[ly + 1][lx + 1];
      const float ex_m2 = __fsub_rn(xm1, xm2), ex_m1 = __fsub_rn(c, xm1), ex_0 = __fsub_rn(xp1, c), ex_p1 = __fsub_rn(xp2, xp1);
      const float ey_m2 = __fsub_rn(ym1, ym2), ey_m1 = __fsub_rn(c, ym1), ey_0 = __fsub_rn(yp1, c), ey_p1 = __fsub_rn(yp2, yp1);
      const float dx2_m2 = __fsub_rn(ex_m1, ex_m2), dx2_m1 = __fsub_rn(ex_0, ex_m1), dx2_0 = __fsub_rn(ex_p1, ex_0);
      const float dy2_m2 = __fsub_rn(ey_m1, ey_m2), dy2_m1 = __fsub_rn(ey_0, ey_m1), dy2_0 = __fsub_rn(ey_p1, ey_0);
      // 2x2 cells touching the centre: dxdy = (D11 - D10) - (D01 - D00) ; dydx = (D11 - D01) - (D10 - D00)
      const float a00 = __fsub_rn(__fsub_rn(pp, yp1), ex_0), b00 = __fsub_rn(__fsub_rn(pp, xp1), ey_0);      // cell (y, x)
      const bool x0ok = x <= w - 3, y0ok = y <= h - 3, c00 = (x <= w - 2) && (y <= h - 2);
      loss += (x0ok ? fabsf(dx2_0) * k_dx2 : 0.f) + (y0ok ? fabsf(dy2_0) * k_dy2 : 0.f) +
              (c00 ? (fabsf(a00) + fabsf(b00)) * k_mix : 0.f);
      if (GRAD) {
        const float a01 = __fsub_rn(__fsub_rn(yp1, pm), ex_m1), b01 = __fsub_rn(ey_0, __fsub_rn(pm, xm1));   // cell (y, x-1)
        const float a10 = __fsub_rn(ex_0, __fsub_rn(mp, ym1)), b10 = __fsub_rn(__fsub_rn(xp1, mp), ey_m1);   // cell (y-1, x)
        const float a11 = __fsub_rn(ex_m1, __fsub_rn(ym1, mm)), b11 = __fsub_rn(ey_m1, __fsub_rn(xm1, mm));  // cell (y-1, x-1)
        float g = 0.f;
        g += (x >= 2) ? sgnc(dx2_m2, k_dx2) : 0.f;
        g -= (x >= 1 && x <= w - 2) ? 2.f * sgnc(dx2_m1, k_dx2) : 0.f;
        g += x0ok ? sgnc(dx2_0, k_dx2) : 0.f;
        g += (y >= 2) ? sgnc(dy2_m2, k_dy2) : 0.f;
        g -= (y >= 1 && y <= h - 2) ? 2.f * sgnc(dy2_m1, k_dy2) : 0.f;
        g += y0ok ? sgnc(dy2_0, k_dy2) : 0.f;
        g += c00 ? (sgnc(a00, k_mix) + sgnc(b00, k_mix)) : 0.f;                                   // centre = D00
        g -= (x >= 1 && y <= h - 2) ? (sgnc(a01, k_mix) + sgnc(b01, k_mix)) : 0.f;                // centre = D01
        g -= (y >= 1 && x <= w - 2) ? (sgnc(a10, k_mix) + sgnc(b10, k_mix)) : 0.f;                // centre = D10
        g += (x >= 1 && y >= 1) ? (sgnc(a11, k_mix) + sgnc(b11, k_mix)) : 0.f;                    // centre = D11
        p.gdisp[s][(size_t)b * h * w + (size_t)y * w + x] = gyv * g;
      }
    }
  }
  loss = sfm_warp_sum(loss);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += s_red[k];
    if (tot != 0.f) atomicAdd(p.acc + 1, (double)tot);
  }
}

}  // namespace

// Fills tiles_x/tiles_y/tile_begin of `p` and launches.  With grad != 0 every gdisp[s] is fully written.
int sfm_launch_smooth(SfmFusedParams& p, int grad, cudaStream_t stream) {
  int total = 0;
  for (int s = 0; s < SFM_MAX_SCALES; ++s) {
    p.tile_begin[s] = total;
    if (s < p.ns) {
      p.tiles_x[s] = (p.w[s] + TS - 1) / TS;
      p.tiles_y[s] = (p.h[s] + TS - 1) / TS;
      total += p.B * p.tiles_x[s] * p.tiles_y[s];
    } else {
      p.tiles_x[s] = p.tiles_y[s] = 1;
    }
  }
  p.tile_begin[SFM_MAX_SCALES] = total;
  if (grad) sfm_smooth_kernel<true><<<total, 256, 0, stream>>>(p);
  else sfm_smooth_kernel<false><<<total, 256, 0, stream>>>(p);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}
