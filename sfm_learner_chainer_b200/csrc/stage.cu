// stage.cu -- stage-level operators behind the reference's function-level API:
//   projective_inverse_warp(imgs, depthes, poses, K)            models/transform.py:156-193
//   SpatialTransformerSamplerInterp forward / backward          models/spational_transformer_sampler_interp.py:32-149
// They work directly on the caller's NCHW arrays (no staging) and exist for callers that use those
// functions on their own and for stage-isolated parity tests; the training hot path is fused_loss.cu.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kT = 256;

struct WarpStageParams {
  int N, h, w;
  const float* imgs;
  const float* depth;
  const float* poses;
  const float* K;
  const float* proj;  // optional override (N,3,4)
  const float* kinv;  // optional override (N,3,3)
  float* out;
  int32_t* u0;
  int32_t* v0;
  uint8_t* inb;
  // backward
  const float* gy;
  float* gdepth;
  float* gposes;
  float* gimgs;
  double* acc;        // N*12
  unsigned* counter;
};

__device__ __forceinline__ void stage_tables(const WarpStageParams& p, int n, float* s_proj, float* s_kinv, float* s_K) {
  if (threadIdx.x == 0) {
    float K[9], pose[6];
    for (int k = 0; k < 9; ++k) K[k] = s_K[k] = p.K[(size_t)n * 9 + k];
    if (p.proj) {
      for (int k = 0; k < 12; ++k) s_proj[k] = p.proj[(size_t)n * 12 + k];
    } else {
      for (int k = 0; k < 6; ++k) pose[k] = p.poses[(size_t)n * 6 + k];
      sfm_build_proj(pose, K, s_proj);
    }
    if (p.kinv) {
      for (int k = 0; k < 9; ++k) s_kinv[k] = p.kinv[(size_t)n * 9 + k];
    } else {
      sfm_inv3(K, s_kinv);
    }
  }
  __syncthreads();
}

struct Taps {
  float a[3], b[3], c[3], d[3];  // I00, I01, I10, I11 per channel
};

__device__ __forceinline__ void load_taps(const float* img, int h, int w, const SfmCoord& c, Taps& t) {
  const size_t plane = (size_t)h * w;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float* pl = img + ch * plane;
    t.a[ch] = c.v00 ? __ldg(pl + (size_t)c.v0 * w + c.u0) : 0.f;
    t.b[ch] = c.v01 ? __ldg(pl + (size_t)c.v0 * w + c.u0 + 1) : 0.f;
    t.c[ch] = c.v10 ? __ldg(pl + (size_t)(c.v0 + 1) * w + c.u0) : 0.f;
    t.d[ch] = c.v11 ? __ldg(pl + (size_t)(c.v0 + 1) * w + c.u0 + 1) : 0.f;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(kT) sfm_warp_stage_kernel(const __grid_constant__ WarpStageParams p) {
  __shared__ float s_proj[12], s_kinv[9], s_K[9];
  __shared__ float s_red[kT / 32][12];
  __shared__ int s_last;
  const int n = blockIdx.y;
  stage_tables(p, n, s_proj, s_kinv, s_K);
  const int h = p.h, w = p.w, hw_n = h * w;
  const int pix = blockIdx.x * kT + threadIdx.x;
  const bool active = pix < hw_n;
  const float hw = (float)((w - 1) / 2.0), hh = (float)((h - 1) / 2.0);
  float dP[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) dP[k] = 0.f;
  if (active) {
    const int y = pix / w, x = pix - y * w;
    const float depth = __ldg(p.depth + (size_t)n * hw_n + pix);
    float rx, ry, rz;
    sfm_ray(s_kinv, (float)x, (float)y, rx, ry, rz);
    const float X = __fmul_rn(depth, rx), Y = __fmul_rn(depth, ry), Z = __fmul_rn(depth, rz);
    SfmCoord c;
    sfm_project(s_proj, X, Y, Z, w, h, hw, hh, c);
    const float* img = p.imgs + (size_t)n * 3 * hw_n;
    Taps t;
    load_taps(img, h, w, c, t);
    const float w1 = __fmul_rn(c.wa, c.wc), w2 = __fmul_rn(c.wb, c.wc);
    const float w3 = __fmul_rn(c.wa, c.wd), w4 = __fmul_rn(c.wb, c.wd);
    if (!BWD) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch)
        p.out[((size_t)n * 3 + ch) * hw_n + pix] = c.any ? sfm_blend(w1, w2, w3, w4, t.a[ch], t.b[ch], t.c[ch], t.d[ch]) : 0.f;
      if (p.u0) p.u0[(size_t)n * hw_n + pix] = c.u0;
      if (p.v0) p.v0[(size_t)n * hw_n + pix] = c.v0;
      if (p.inb) p.inb[(size_t)n * hw_n + pix] = c.inb ? 1 : 0;
    } else {
      float g[3], gu = 0.f, gv = 0.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        g[ch] = __ldg(p.gy + ((size_t)n * 3 + ch) * hw_n + pix);
        gu += g[ch] * (c.wc * (t.b[ch] - t.a[ch]) + c.wd * (t.d[ch] - t.c[ch]));
        gv += g[ch] * (c.wa * (t.c[ch] - t.a[ch]) + c.wb * (t.d[ch] - t.b[ch]));
      }
      const float rzv = 1.f / c.z;
      float gq0 = gu * c.fx * rzv, gq1 = gv * c.fy * rzv;
      float gq2 = -(gq0 * c.q0 + gq1 * c.q1) * rzv;
      if (!c.any) { gq0 = 0.f; gq1 = 0.f; gq2 = 0.f; }
      const float gX = gq0 * s_proj[0] + gq1 * s_proj[4] + gq2 * s_proj[8];
      const float gY = gq0 * s_proj[1] + gq1 * s_proj[5] + gq2 * s_proj[9];
      const float gZ = gq0 * s_proj[2] + gq1 * s_proj[6] + gq2 * s_proj[10];
      p.gdepth[(size_t)n * hw_n + pix] = gX * rx + gY * ry + gZ * rz;
      dP[0] = gq0 * X; dP[1] = gq0 * Y; dP[2] = gq0 * Z; dP[3] = gq0;
      dP[4] = gq1 * X; dP[5] = gq1 * Y; dP[6] = gq1 * Z; dP[7] = gq1;
      dP[8] = gq2 * X; dP[9] = gq2 * Y; dP[10] = gq2 * Z; dP[11] = gq2;
      if (p.gimgs) {
        float* gi = p.gimgs + (size_t)n * 3 * hw_n;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float* pl = gi + (size_t)ch * hw_n;
          if (c.v00) atomicAdd(pl + (size_t)c.v0 * w + c.u0, g[ch] * w1);
          if (c.v01) atomicAdd(pl + (size_t)c.v0 * w + c.u0 + 1, g[ch] * w2);
          if (c.v10) atomicAdd(pl + (size_t)(c.v0 + 1) * w + c.u0, g[ch] * w3);
          if (c.v11) atomicAdd(pl + (size_t)(c.v0 + 1) * w + c.u0 + 1, g[ch] * w4);
        }
      }
    }
  }
  if (!BWD) return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float v = sfm_warp_sum(dP[k]);
    if (lane == 0) s_red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float a = 0.f;
    for (int k = 0; k < kT / 32; ++k) a += s_red[k][threadIdx.x];
    s_red[0][threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    const int rr = threadIdx.x / 4, j = threadIdx.x % 4;
    const float v = s_K[0 * 3 + rr] * s_red[0][0 * 4 + j] + s_K[1 * 3 + rr] * s_red[0][1 * 4 + j] +
                    s_K[2 * 3 + rr] * s_red[0][2 * 4 + j];
    atomicAdd(p.acc + (size_t)n * 12 + threadIdx.x, (double)v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(p.counter, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < p.N; e += kT) {
    double dT[12];
    float pose[6], g[6];
    for (int k = 0; k < 12; ++k) dT[k] = __ldcg(p.acc + (size_t)e * 12 + k);
    for (int k = 0; k < 6; ++k) pose[k] = p.poses[(size_t)e * 6 + k];
    sfm_pose_backward(pose, dT, g);
    for (int k = 0; k < 6; ++k) p.gposes[(size_t)e * 6 + k] = g[k];
  }
}

// ---- SpatialTransformerSamplerInterp ---------------------------------------------------------
struct InterpCoord {
  int u0, u1, v0, v1;
  float wx0, wx1, wy0, wy1;
};

__device__ __forceinline__ InterpCoord interp_coord(float u, float v, int H, int W) {
  // floor, +1, clamp to the image, weights from the CLAMPED indices (interp.py:41-55)
  float u0 = floorf(u), v0 = floorf(v);
  float u1 = __fadd_rn(u0, 1.f), v1 = __fadd_rn(v0, 1.f);
  u0 = fminf(fmaxf(u0, 0.f), (float)(W - 1));
  v0 = fminf(fmaxf(v0, 0.f), (float)(H - 1));
  u1 = fminf(fmaxf(u1, 0.f), (float)(W - 1));
  v1 = fminf(fmaxf(v1, 0.f), (float)(H - 1));
  InterpCoord c;
  c.wx0 = __fsub_rn(u1, u);
  c.wx1 = __fsub_rn(u, u0);
  c.wy0 = __fsub_rn(v1, v);
  c.wy1 = __fsub_rn(v, v0);
  c.u0 = (int)u0; c.u1 = (int)u1; c.v0 = (int)v0; c.v1 = (int)v1;
  return c;
}

__global__ void sfm_interp_fwd_kernel(int B, int C, int H, int W, int oH, int oW, const float* __restrict__ x,
                                      const float* __restrict__ grid, float* __restrict__ y) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int on = oH * oW;
  if (gid >= (long long)B * on) return;
  const int b = (int)(gid / on), pix = (int)(gid - (long long)b * on);
  const float u = grid[((size_t)b * 2 + 0) * on + pix], v = grid[((size_t)b * 2 + 1) * on + pix];
  const InterpCoord c = interp_coord(u, v, H, W);
  const float w1 = __fmul_rn(c.wx0, c.wy0), w2 = __fmul_rn(c.wx1, c.wy0);
  const float w3 = __fmul_rn(c.wx0, c.wy1), w4 = __fmul_rn(c.wx1, c.wy1);
  for (int ch = 0; ch < C; ++ch) {
    const float* pl = x + ((size_t)b * C + ch) * H * W;
    float acc = __fmul_rn(w1, pl[(size_t)c.v0 * W + c.u0]);
    acc = __fadd_rn(acc, __fmul_rn(w2, pl[(size_t)c.v0 * W + c.u1]));
    acc = __fadd_rn(acc, __fmul_rn(w3, pl[(size_t)c.v1 * W + c.u0]));
    acc = __fadd_rn(acc, __fmul_rn(w4, pl[(size_t)c.v1 * W + c.u1]));
    y[((size_t)b * C + ch) * on + pix] = acc;
  }
}

__global__ void sfm_interp_bwd_kernel(int B, int C, int H, int W, int oH, int oW, const float* __restrict__ x,
                                      const float* __restrict__ grid, const float* __restrict__ gy,
                                      float* __restrict__ ggrid) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int on = oH * oW;
  if (gid >= (long long)B * on) return;
  const int b = (int)(gid / on), pix = (int)(gid - (long long)b * on);
  const float u = grid[((size_t)b * 2 + 0) * on + pix], v = grid[((size_t)b * 2 + 1) * on + pix];
  const InterpCoord c = interp_coord(u, v, H, W);
  float gu = 0.f, gv = 0.f;
  for (int ch = 0; ch < C; ++ch) {
    const float* pl = x + ((size_t)b * C + ch) * H * W;
    const float x1 = pl[(size_t)c.v0 * W + c.u0], x2 = pl[(size_t)c.v0 * W + c.u1];
    const float x3 = pl[(size_t)c.v1 * W + c.u0], x4 = pl[(size_t)c.v1 * W + c.u1];
    float a = __fmul_rn(-c.wy0, x1);
    a = __fadd_rn(a, __fmul_rn(c.wy0, x2));
    a = __fsub_rn(a, __fmul_rn(c.wy1, x3));
    a = __fadd_rn(a, __fmul_rn(c.wy1, x4));
    float bb = __fmul_rn(-c.wx0, x1);
    bb = __fsub_rn(bb, __fmul_rn(c.wx1, x2));
    bb = __fadd_rn(bb, __fmul_rn(c.wx0, x3));
    bb = __fadd_rn(bb, __fmul_rn(c.wx1, x4));
    const float g = gy[((size_t)b * C + ch) * on + pix];
    gu = __fadd_rn(gu, __fmul_rn(a, g));
    gv = __fadd_rn(gv, __fmul_rn(bb, g));
  }
  ggrid[((size_t)b * 2 + 0) * on + pix] = gu;
  ggrid[((size_t)b * 2 + 1) * on + pix] = gv;
}

}  // namespace

extern "C" int sfm_warp_forward(int N, int h, int w, const float* imgs, const float* depth, const float* poses,
                                const float* K, const float* proj, const float* kinv, float* out, int32_t* u0,
                                int32_t* v0, uint8_t* inb, void* stream) {
  if (N <= 0 || h < 1 || w < 1) { sfm_set_error("sfm_warp_forward: invalid shape N=%d h=%d w=%d", N, h, w); return SFM_E_INVALID_SHAPE; }
  if (!imgs || !depth || !K || !out || (!poses && !proj)) { sfm_set_error("sfm_warp_forward: null pointer"); return SFM_E_NULL_POINTER; }
  WarpStageParams p{};
  p.N = N; p.h = h; p.w = w; p.imgs = imgs; p.depth = depth; p.poses = poses; p.K = K; p.proj = proj; p.kinv = kinv;
  p.out = out; p.u0 = u0; p.v0 = v0; p.inb = inb;
  dim3 grid((h * w + kT - 1) / kT, N);
  sfm_warp_stage_kernel<false><<<grid, kT, 0, (cudaStream_t)stream>>>(p);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" size_t sfm_warp_backward_scratch_bytes(int N) { return (size_t)(N > 0 ? N : 0) * 12 * sizeof(double) + 16; }

extern "C" int sfm_warp_backward(int N, int h, int w, const float* imgs, const float* depth, const float* poses,
                                 const float* K, const float* gy, float* gdepth, float* gposes, float* gimgs,
                                 void* scratch, void* stream) {
  if (N <= 0 || h < 1 || w < 1) { sfm_set_error("sfm_warp_backward: invalid shape N=%d h=%d w=%d", N, h, w); return SFM_E_INVALID_SHAPE; }
  if (!imgs || !depth || !K || !poses || !gy || !gdepth || !gposes || !scratch) { sfm_set_error("sfm_warp_backward: null pointer"); return SFM_E_NULL_POINTER; }
  cudaStream_t st = (cudaStream_t)stream;
  SFM_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sfm_warp_backward_scratch_bytes(N), st));
  if (gimgs) SFM_CUDA_CHECK(cudaMemsetAsync(gimgs, 0, (size_t)N * 3 * h * w * sizeof(float), st));
  WarpStageParams p{};
  p.N = N; p.h = h; p.w = w; p.imgs = imgs; p.depth = depth; p.poses = poses; p.K = K;
  p.gy = gy; p.gdepth = gdepth; p.gposes = gposes; p.gimgs = gimgs;
  p.acc = (double*)scratch;
  p.counter = (unsigned*)((char*)scratch + (size_t)N * 12 * sizeof(double));
  dim3 grid((h * w + kT - 1) / kT, N);
  sfm_warp_stage_kernel<true><<<grid, kT, 0, st>>>(p);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int sfm_sampler_interp_forward(int B, int C, int H, int W, int oH, int oW, const float* x, const float* grid,
                                          float* y, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || oH <= 0 || oW <= 0) { sfm_set_error("sfm_sampler_interp_forward: invalid shape"); return SFM_E_INVALID_SHAPE; }
  if (!x || !grid || !y) { sfm_set_error("sfm_sampler_interp_forward: null pointer"); return SFM_E_NULL_POINTER; }
  const long long n = (long long)B * oH * oW;
  sfm_interp_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, C, H, W, oH, oW, x, grid, y);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int sfm_sampler_interp_backward(int B, int C, int H, int W, int oH, int oW, const float* x,
                                           const float* grid, const float* gy, float* gx, float* ggrid, void* stream) {
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || oH <= 0 || oW <= 0) { sfm_set_error("sfm_sampler_interp_backward: invalid shape"); return SFM_E_INVALID_SHAPE; }
  if (!x || !grid || !gy || !ggrid) { sfm_set_error("sfm_sampler_interp_backward: null pointer"); return SFM_E_NULL_POINTER; }
  cudaStream_t st = (cudaStream_t)stream;
  if (gx) SFM_CUDA_CHECK(cudaMemsetAsync(gx, 0, (size_t)B * C * H * W * sizeof(float), st));   // gx = zeros (interp.py:148)
  const long long n = (long long)B * oH * oW;
  sfm_interp_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(B, C, H, W, oH, oW, x, grid, gy, ggrid);
  SFM_CUDA_CHECK(cudaGetLastError());
  return 0;
}
