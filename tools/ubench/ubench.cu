// ubench.cu -- instruction-throughput microbenchmarks for sm_100a (B200), used to size the fused loss kernels.
// Each test runs NW warps per SMSP on every SM, each thread executing ITERS x (UNROLL independent chains) of one op
// class, and reports warp-instructions per clock per SMSP from clock64() deltas.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;
constexpr int U = 8;

struct OpFFMA  { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = __fmaf_rn(a, b, c); } };
struct OpFMUL  { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = __fmul_rn(a, b); } };
struct OpFADD  { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = __fadd_rn(a, c); } };
struct OpMULADD{ static constexpr int n = 2; __device__ static void op(float& a, float b, float c) { a = __fadd_rn(__fmul_rn(a, b), c); } };
struct OpFMNMX { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = fminf(a, c); } };
struct OpRCP   { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); } };
struct OpEX2   { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); } };
struct OpFDIV  { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = __fdiv_rn(a, b); } };
struct OpFRCPRN{ static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = __frcp_rn(a); } };
struct OpFLOOR { static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = floorf(a); asm volatile("" : "+f"(a)); } };
struct OpF2I2F { static constexpr int n = 2; __device__ static void op(float& a, float b, float c) { int i = __float2int_rd(a); asm volatile("" : "+r"(i)); a = (float)i; } };
struct OpFADDRM{ static constexpr int n = 1; __device__ static void op(float& a, float b, float c) { a = __fadd_rd(a, c); } };
struct OpSETSEL{ static constexpr int n = 2; __device__ static void op(float& a, float b, float c) { a = (a > b) ? c : a; asm volatile("" : "+f"(a)); } };
struct OpFFMAIADD { static constexpr int n = 2; __device__ static void op(float& a, float b, float c) {
  a = __fmaf_rn(a, b, c); int i = __float_as_int(c); asm volatile("" : "+r"(i)); } };

template <class OP>
__global__ void k_scalar(float* out, long long* cyc, float b, float c) {
  float a[U];
#pragma unroll
  for (int j = 0; j < U; ++j) a[j] = threadIdx.x * 0.001f + j + 1.f;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < U; ++j) OP::op(a[j], b, c);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < U; ++j) s += a[j];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// int ALU ops mixed with FFMA: tests dual issue of the fma and alu pipes
template <int MODE>
__global__ void k_mix(float* out, long long* cyc, float b, float c, int ib) {
  float a[U]; int q[U];
#pragma unroll
  for (int j = 0; j < U; ++j) { a[j] = threadIdx.x * 0.001f + j + 1.f; q[j] = threadIdx.x + j; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (MODE == 0) { q[j] = (q[j] ^ ib) + it; }                       // LOP3 + IADD3 (alu only)
      if (MODE == 1) { a[j] = __fmaf_rn(a[j], b, c); q[j] = (q[j] & ib) ^ it; }   // FFMA + LOP3
      if (MODE == 2) { a[j] = __fmaf_rn(a[j], b, c); a[j] = fminf(a[j], c); }    // FFMA + FMNMX (dependent)
      if (MODE == 3) { a[j] = __fadd_rn(a[j], c); q[j] = (q[j] & ib) ^ it; }      // FADD + LOP3
      if (MODE == 4) { a[j] = __fmaf_rn(a[j], b, c); q[j] = q[j] * ib + it; }     // FFMA + IMAD (same pipe?)
    }
  }
  long long t1 = clock64();
  float s = 0; int qs = 0;
#pragma unroll
  for (int j = 0; j < U; ++j) { s += a[j]; qs += q[j]; }
  if (s == 123.456f || qs == 0x7fffffff) out[0] = s + qs;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
__global__ void k_packed(float* out, long long* cyc, float b, float c) {
  float2 a[U];
  const float2 bb = make_float2(b, b * 1.0001f), cc = make_float2(c, c * 1.0001f);
#pragma unroll
  for (int j = 0; j < U; ++j) a[j] = make_float2(threadIdx.x * 0.001f + j + 1.f, threadIdx.x * 0.002f + j);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (MODE == 0) a[j] = __ffma2_rn(a[j], bb, cc);
      if (MODE == 1) a[j] = __fmul2_rn(a[j], bb);
      if (MODE == 2) a[j] = __fadd2_rn(a[j], cc);
      if (MODE == 3) a[j] = __fadd2_rd(a[j], cc);
      if (MODE == 4) a[j] = __fadd2_rn(__fmul2_rn(a[j], bb), cc);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < U; ++j) s += a[j].x + a[j].y;
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shuffle / shared-memory throughput
template <int MODE>
__global__ void k_shfl(float* out, long long* cyc) {
  __shared__ float4 sm[1024];
  float a[U];
#pragma unroll
  for (int j = 0; j < U; ++j) a[j] = threadIdx.x + j;
  sm[threadIdx.x] = make_float4(1, 2, 3, 4);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (MODE == 0) a[j] = __shfl_up_sync(0xffffffffu, a[j], 1);
      if (MODE == 1) { float4 v = sm[(threadIdx.x + j * 32 + it) & 1023]; a[j] += v.x + v.w; }
      if (MODE == 2) { float v = ((float*)sm)[(threadIdx.x + j * 32 + it) & 4095]; a[j] += v; }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < U; ++j) s += a[j];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Gather: every lane loads 4 float4 texels (2x2 bilinear footprint) around a smoothly varying position, like the
// warp sampler.  footprint_bytes selects L1 / L2 / HBM residency.
__global__ void k_gather(const float4* __restrict__ img, int w, int h, int nimg, float4* out, long long* cyc, int iters, float shift) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  // warp marches down rows of strip (warp % strips) of image (warp / strips)
  const int strips = w / 32;
  for (int it = 0; it < iters; ++it) {
    const int task = warp + it * nwarps;
    const int im = (task / strips / (h / 16)) % nimg;
    const int strip = task % strips;
    const int seg = (task / strips) % (h / 16);
#pragma unroll 4
    for (int r = 0; r < 16; ++r) {
      const int y = seg * 16 + r;
      const float u = fminf(fmaxf(strip * 32 + lane * 1.03f + shift, 0.f), w - 2.f);
      const float v = fminf(fmaxf(y * 1.01f + shift * 0.3f, 0.f), h - 2.f);
      const int u0 = (int)u, v0 = (int)v;
      const float4* p = img + ((size_t)im * h + v0) * w + u0;
      const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + w), d = __ldg(p + w + 1);
      acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y; acc.z += a.z + b.z + c.z + d.z;
    }
  }
  long long t1 = clock64();
  if (acc.x == 123.456f) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static int g_sms = 148;

template <class F>
void run(const char* name, int instr_per_op, int warps_per_smsp, F launch) {
  const int blocks = g_sms, threads = warps_per_smsp * 4 * 32;
  long long* cyc; float* out;
  CK(cudaMalloc(&cyc, blocks * sizeof(long long))); CK(cudaMalloc(&out, 64));
  launch(blocks, threads, out, cyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(blocks, threads, out, cyc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(blocks);
  CK(cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  const double med = (double)h[blocks / 2];
  const double winstr = (double)ITERS * U * instr_per_op * warps_per_smsp;   // per SMSP
  printf("%-28s nw/smsp=%2d  cycles=%9.0f  warp-instr/clk/SMSP=%6.3f  (%.3f ms, %.2f GHz eff)\n", name, warps_per_smsp, med,
         winstr / med, ms, med / (ms * 1e6));
  cudaFree(cyc); cudaFree(out);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  g_sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %d MB\n", prop.name, g_sms, prop.l2CacheSize >> 20);
  for (int nw : {1, 2, 4, 8}) {
#define RS(OP) run(#OP, OP::n, nw, [&](int b, int t, float* o, long long* c) { k_scalar<OP><<<b, t>>>(o, c, 1.0001f, 0.5f); })
    RS(OpFFMA); RS(OpFMUL); RS(OpFADD); RS(OpMULADD); RS(OpFMNMX); RS(OpRCP); RS(OpEX2); RS(OpFDIV); RS(OpFRCPRN);
    RS(OpFLOOR); RS(OpF2I2F); RS(OpFADDRM); RS(OpSETSEL);
    run("LOP3+IADD3", 2, nw, [&](int b, int t, float* o, long long* c) { k_mix<0><<<b, t>>>(o, c, 1.0001f, 0.5f, 0x55); });
    run("FFMA+LOP3x2", 3, nw, [&](int b, int t, float* o, long long* c) { k_mix<1><<<b, t>>>(o, c, 1.0001f, 0.5f, 0x55); });
    run("FFMA+FMNMX", 2, nw, [&](int b, int t, float* o, long long* c) { k_mix<2><<<b, t>>>(o, c, 1.0001f, 0.5f, 0x55); });
    run("FADD+LOP3x2", 3, nw, [&](int b, int t, float* o, long long* c) { k_mix<3><<<b, t>>>(o, c, 1.0001f, 0.5f, 0x55); });
    run("FFMA+IMAD", 2, nw, [&](int b, int t, float* o, long long* c) { k_mix<4><<<b, t>>>(o, c, 1.0001f, 0.5f, 0x55); });
    run("FFMA2", 1, nw, [&](int b, int t, float* o, long long* c) { k_packed<0><<<b, t>>>(o, c, 1.0001f, 0.5f); });
    run("FMUL2", 1, nw, [&](int b, int t, float* o, long long* c) { k_packed<1><<<b, t>>>(o, c, 1.0001f, 0.5f); });
    run("FADD2", 1, nw, [&](int b, int t, float* o, long long* c) { k_packed<2><<<b, t>>>(o, c, 1.0001f, 0.5f); });
    run("FADD2.RM", 1, nw, [&](int b, int t, float* o, long long* c) { k_packed<3><<<b, t>>>(o, c, 1.0001f, 0.5f); });
    run("FMUL2+FADD2", 2, nw, [&](int b, int t, float* o, long long* c) { k_packed<4><<<b, t>>>(o, c, 1.0001f, 0.5f); });
    run("SHFL.UP", 1, nw, [&](int b, int t, float* o, long long* c) { k_shfl<0><<<b, t>>>(o, c); });
    run("LDS.128(+2 FADD)", 3, nw, [&](int b, int t, float* o, long long* c) { k_shfl<1><<<b, t>>>(o, c); });
    run("LDS.32(+FADD)", 2, nw, [&](int b, int t, float* o, long long* c) { k_shfl<2><<<b, t>>>(o, c); });
    printf("\n");
  }
  // gather test: images of 128x416 float4 texels (852 KB each)
  {
    const int w = 416, h = 128;
    for (int nimg : {8, 64, 1024}) {
      float4* img; CK(cudaMalloc(&img, (size_t)nimg * w * h * 16)); CK(cudaMemset(img, 0, (size_t)nimg * w * h * 16));
      float4* out; long long* cyc; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 148 * 64 * 8));
      for (int wpb : {1, 4}) for (int bps : {8, 16, 32}) {
        if (wpb == 4 && bps > 8) continue;
        const int blocks = g_sms * bps, threads = 32 * wpb;
        const int nwarps = blocks * wpb;
        const int tasks = nimg * (w / 32) * (h / 16);
        const int iters = std::max(1, tasks / nwarps);
        k_gather<<<blocks, threads>>>(img, w, h, nimg, out, cyc, iters, 3.3f);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_gather<<<blocks, threads>>>(img, w, h, nimg, out, cyc, iters, 3.3f);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double px = (double)nwarps * iters * 16 * 32;
        printf("gather nimg=%4d (%6.1f MB) blocks/SM=%2d warps/blk=%d: %.3f ms, %.1f Gpix-src/s, %.1f GB/s unique(16B/px)\n", nimg,
               nimg * w * h * 16 / 1e6, bps, wpb, ms, px / ms / 1e6, px * 16 / ms / 1e6);
      }
      cudaFree(img); cudaFree(out); cudaFree(cyc);
    }
  }
  return 0;
}
