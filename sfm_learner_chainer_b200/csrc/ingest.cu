// ingest.cu -- the data layer in front of the loss path, fused into one gather kernel:
//   load_as_float_norm           datasets/kitti/kitti_raw_dataset.py:12-14   uint8 HWC frame -> float32 CHW, x / 127.5 - 1
//   data_augmentation            datasets/kitti/kitti_raw_transformed.py:23-74 random scaling (F.resize_images to
//                                (int(H*sy), int(W*sx)), :32-44), random crop back to (H, W) (:47-58), random
//                                horizontal flip (:61-66), each with its update of the intrinsics
//   get_multi_scale_intrinsics   :76-93                                       K / 2^s pyramid
// The random draws stay with the caller (its RNG, its order: scaling, offsets, flip); the kernel takes them as one
// SfmAugment per snippet.  Every output pixel is traced back through flip -> crop -> resize to four uint8 taps of
// the decoded frame, so the host never touches float images: the H2D copy carries 1 byte per sample instead of
// 4 and the dataset workers' CPU resize disappears.
//
// Arithmetic (bit-exact against oracle/sfm_oracle.py:ingest_u8): taps are normalised individually,
// n = fp32(p) / 127.5f - 1.f, and blended as Chainer's resize_images does (float64 linspace coordinates, last one
// pinned; float64 weight products cast to fp32; y = ((w1 a + w2 b) + w3 c) + w4 d) -- the same code as the
// pyramid kernel.  Intrinsics: fx' = fp32(fp64(fx) * sx) (numpy float32 scalar times float64 scalar), cx' likewise,
// then cx' - off_x, then W - cx' when flipped, then / 2^s.
#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ float norm_u8(uint8_t p) { return __fsub_rn(__fdiv_rn((float)p, 127.5f), 1.f); }

// 256-entry table of norm_u8 in shared memory: the IEEE division costs ~17 issue slots and every output pixel needs
// twelve of them; the table holds exactly the values the division would produce.
__global__ void __launch_bounds__(256) sfm_ingest_u8_kernel(const uint8_t* __restrict__ frames, const SfmAugment* __restrict__ aug,
                                                            float* __restrict__ tgt, float* __restrict__ src, int B, int S, int H,
                                                            int W) {
  __shared__ float s_norm[256];
  s_norm[threadIdx.x] = norm_u8((uint8_t)threadIdx.x);
  __syncthreads();
  const int img = blockIdx.y;                 // b * (1 + S) + j ; j == 0 is the target frame
  const int b = img / (1 + S), j = img - b * (1 + S);
  SfmAugment a;
  if (aug) a = aug[b];
  else { a.out_h = H; a.out_w = W; a.off_y = a.off_x = a.flip = 0; a.x_scaling = a.y_scaling = 1.0; }
  const size_t plane = (size_t)H * W;
  const uint8_t* __restrict__ in = frames + (size_t)img * plane * 3;
  float* __restrict__ out = (j == 0) ? tgt + (size_t)b * 3 * plane : src + ((size_t)b * S + (j - 1)) * 3 * plane;
  const double stepx = (a.out_w > 1) ? __ddiv_rn((double)(W - 1), (double)(a.out_w - 1)) : 0.0;
  const double stepy = (a.out_h > 1) ? __ddiv_rn((double)(H - 1), (double)(a.out_h - 1)) : 0.0;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < H * W; pix += gridDim.x * blockDim.x) {
    const int y = pix / W, x = pix - y * W;
    const int X = a.off_x + (a.flip ? W - 1 - x : x);          // column / row in the rescaled image
    const int Y = a.off_y + y;
    const double u = (X == a.out_w - 1 && a.out_w > 1) ? (double)(W - 1) : __dmul_rn((double)X, stepx);
    const double v = (Y == a.out_h - 1 && a.out_h > 1) ? (double)(H - 1) : __dmul_rn((double)Y, stepy);
    const int u0 = min(max((int)floor(u), 0), W - 2), v0 = min(max((int)floor(v), 0), H - 2);
    const double ua = __dsub_rn((double)(u0 + 1), u), ub = __dsub_rn(u, (double)u0);
    const double va = __dsub_rn((double)(v0 + 1), v), vb = __dsub_rn(v, (double)v0);
    const float w1 = (float)__dmul_rn(va, ua), w2 = (float)__dmul_rn(va, ub);
    const float w3 = (float)__dmul_rn(vb, ua), w4 = (float)__dmul_rn(vb, ub);
    const uint8_t* __restrict__ t0 = in + ((size_t)v0 * W + u0) * 3;      // taps (v0,u0),(v0,u0+1): 6 contiguous bytes
    const uint8_t* __restrict__ t1 = t0 + (size_t)W * 3;
    uint8_t q[12];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      q[k] = __ldg(t0 + k);
      q[6 + k] = __ldg(t1 + k);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[c * plane + pix] = sfm_blend(w1, w2, w3, w4, s_norm[q[c]], s_norm[q[3 + c]], s_norm[q[6 + c]], s_norm[q[9 + c]]);
  }
}

__global__ void sfm_ingest_intrinsics_kernel(const float* __restrict__ K_in, const SfmAugment* __restrict__ aug,
                                             float* __restrict__ K_out, int B, int W, int ns) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* K = K_in + (size_t)b * 9;
  float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  if (aug) {
    const SfmAugment a = aug[b];
    fx = (float)__dmul_rn((double)fx, a.x_scaling);      // kitti_raw_transformed.py:39-42
    fy = (float)__dmul_rn((double)fy, a.y_scaling);
    cx = (float)__dmul_rn((double)cx, a.x_scaling);
    cy = (float)__dmul_rn((double)cy, a.y_scaling);
    cx = __fsub_rn(cx, (float)a.off_x);                  // :53-54
    cy = __fsub_rn(cy, (float)a.off_y);
    if (a.flip) cx = __fsub_rn((float)W, cx);            // :65
  }
  for (int s = 0; s < ns; ++s) {                         // :86-91
    const float d = (float)(1 << s);
    float* o = K_out + ((size_t)b * ns + s) * 9;
    o[0] = __fdiv_rn(fx, d); o[1] = 0.f; o[2] = __fdiv_rn(cx, d);
    o[3] = 0.f; o[4] = __fdiv_rn(fy, d); o[5] = __fdiv_rn(cy, d);
    o[6] = 0.f; o[7] = 0.f; o[8] = 1.f;
  }
}

}  // namespace

int sfm_launch_ingest_u8(int B, int S, int H, int W, int ns, const uint8_t* frames, const float* K_in, const SfmAugment* aug,
                         float* tgt, float* src, float* K_out, cudaStream_t stream) {
  const int per_img = (H * W + 255) / 256;
  dim3 grid((unsigned)(per_img > 64 ? 64 : per_img), (unsigned)(B * (1 + S)));
  sfm_ingest_u8_kernel<<<grid, 256, 0, stream>>>(frames, aug, tgt, src, B, S, H, W);
  SFM_CUDA_CHECK(cudaGetLastError());
  if (K_out) {
    sfm_ingest_intrinsics_kernel<<<(B + 127) / 128, 128, 0, stream>>>(K_in, aug, K_out, B, W, ns);
    SFM_CUDA_CHECK(cudaGetLastError());
  }
  return 0;
}
