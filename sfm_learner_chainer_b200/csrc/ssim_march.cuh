// ssim_march.cuh -- SSIM marching kernel: L1 + SSIM photometric terms (base_model.py:110-115, 126-142),
// forward and backward in one march.  Included by fused_loss.cu (shares its projection / blend helpers).
//
// A warp owns a strip of 28 interior columns (+2 halo columns each side = 32 lanes) x hseg rows and
// marches down rows y0-2 .. y1+1, lane = column.  Per row r:
//   stage A  warp the pixel (r, lane): projection, 4-tap gather, blend, image gradients dP/du, dP/dv
//   stage B  horizontal 3-sums of P, P^2, P.T, T, T^2 (neighbours through warp shuffles)
//   stage C  vertical 3-sums (register ring over the last three rows) -> SSIM at (r-1, lane), its loss
//            and the three gradient fields g_a, g_s, g_c
//   stage D  horizontal 3-sums of the gradient fields
//   stage E  vertical 3-sums -> dL/dP at (r-2, lane) = A(g_a) + 2P.A(g_s) + T.A(g_c) + L1 part
//   stage F  sampler / projection backward of pixel (r-2, lane) from its forward record
// The rings and the two-row delay line of forward records live in registers; the row loop is unrolled
// three times so that the ring rotation is pure renaming.  No shared memory, no barrier; the only
// atomics are the per-task flush.  dL/dP accumulates as sum gq, sum gq*depth, sum gq*depth*y per lane
// (the lane's column x is constant along the march) and is expanded with Kinv before the flush.
#pragma once

namespace {

constexpr int SSIM_IW = 28;     // interior columns per strip

// Three channels as one packed pair (channels 0, 1: FFMA2 / FADD2 / FMUL2 of sm_100, one issue slot for two
// lanes) plus a scalar (channel 2).  Used for the SSIM statistics and gradient fields (tolerance-checked region; the
// individually rounded coordinate chain and the blend stay scalar).  The packed forms have the same FP32 throughput
// as scalar code; what they save is issue slots, which is what bounds this kernel.
struct V3 {
  float2 a;
  float b;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.a = make_float2(x, y); r.b = z; return r; }
__device__ __forceinline__ V3 v3s(float s) { return v3(s, s, s); }
__device__ __forceinline__ V3 operator+(const V3& x, const V3& y) { V3 r; r.a = __fadd2_rn(x.a, y.a); r.b = x.b + y.b; return r; }
__device__ __forceinline__ V3 operator*(const V3& x, const V3& y) { V3 r; r.a = __fmul2_rn(x.a, y.a); r.b = x.b * y.b; return r; }
__device__ __forceinline__ V3 vfma(const V3& x, const V3& y, const V3& z) { V3 r; r.a = __ffma2_rn(x.a, y.a, z.a); r.b = fmaf(x.b, y.b, z.b); return r; }
__device__ __forceinline__ V3 vneg(const V3& x) { return v3(-x.a.x, -x.a.y, -x.b); }
__device__ __forceinline__ V3 vshfl_up(const V3& x) {
  return v3(__shfl_up_sync(0xffffffffu, x.a.x, 1), __shfl_up_sync(0xffffffffu, x.a.y, 1), __shfl_up_sync(0xffffffffu, x.b, 1));
}
__device__ __forceinline__ V3 vshfl_down(const V3& x) {
  return v3(__shfl_down_sync(0xffffffffu, x.a.x, 1), __shfl_down_sync(0xffffffffu, x.a.y, 1), __shfl_down_sync(0xffffffffu, x.b, 1));
}

struct StripTask {
  int s, b, x0, y0, y1;
};

__device__ __forceinline__ StripTask decode_strip(const SfmFusedParams& p, int t) {
  StripTask k;
  int s = 0;
#pragma unroll
  for (int q = 1; q < SFM_MAX_SCALES; ++q)
    if (q < p.ns && t >= p.task_begin[q]) s = q;
  t -= p.task_begin[s];
  k.s = s;
  const int seg = t % p.nseg[s];
  t /= p.nseg[s];
  const int strip = t % p.nstrip[s];
  k.b = t / p.nstrip[s];
  k.x0 = strip * SSIM_IW;
  k.y0 = seg * p.hseg;
  k.y1 = min(k.y0 + p.hseg, p.h[s]);
  return k;
}

// forward record of one pixel, kept for two rows until its backward runs
struct Rec {
  float Ix[3], Iy[3];     // dP_c/du, dP_c/dv (pixel units)
  float P[3], T[3];
  float q0, q1, q2, r;    // projection, 1/z (0 when out of view)
  float depth;
  float dsc;              // depth, or depth * (d disp / d x) when the disparity input is the pre-activation map
};

template <int N>
struct IC { static constexpr int value = N; };

#ifndef SFM_SSIM_FENCE
#define SFM_SSIM_FENCE 1      // basic-block fence after the refill (see step)
#endif
#ifndef SFM_MINB_SSIM
#define SFM_MINB_SSIM 12      // resident warps per SM of the instances that keep the forward records in shared memory (168 registers)
#endif
#ifndef SFM_MINB_SSIM_REG
#define SFM_MINB_SSIM_REG 8   // ... of the instances that keep them in registers (250 registers)
#endif
// NW = warps per CTA.  NW == 1: one warp walks the strip segment once per source.  NW == 2 (two sources, small
// batches): the two warps of a CTA walk the same segment at the same time, one source each, so a task is half as
// long and -- for the same number of resident warps -- twice as tall (half the halo rows).  gdisp stays
// deterministic: warp 1 hands its per-row term to warp 0 through shared memory (one CTA barrier per row) and warp 0
// applies both in the order of the sequential loop, with the same fused multiply-adds.
// SREC: where the two-row delay line of forward records lives.  true: shared memory (34 LDS/STS wavefronts per row,
// 168 registers, 3 warps per scheduler) -- the throughput form for grids of several waves.  false: registers (250
// registers, 2 warps per scheduler, 6 % fewer instructions and a third fewer L1 data-pipe wavefronts) -- faster whenever
// the whole grid is resident at once anyway (cfg2: 29.3 -> 26.6 us), slower when occupancy counts (cfg5: 933 -> 952 us).
template <bool GRAD, bool ACCUM, bool DEBUG, bool RAW, int NW, bool SREC = true>
__global__ void __launch_bounds__(32 * NW, (SREC ? SFM_MINB_SSIM : SFM_MINB_SSIM_REG) / NW)
sfm_ssim_march_kernel(const __grid_constant__ SfmFusedParams p) {
  constexpr bool SFM_SSIM_SREC = SREC;
  __shared__ float4 sP_all[NW][3];
  __shared__ float sG[(NW > 1) ? 3 * 2 * 32 : 1];       // [ring slot][gdd | dsc][lane] of warp 1's row term
  const int wi = (NW > 1) ? (int)(threadIdx.x >> 5) : 0;
  float4* const sP = sP_all[wi];
  // warp-uniform base pointers, re-read (broadcast LDS.64) where they are used: at 168 registers ptxas otherwise
  // re-derives them from the kernel parameters inside the row loop (64-bit multiplies per row, seen in the SASS)
  struct RowPtrs {
    const float* img;
    const float* tgt;
    const float* disp;
    float* gdisp;
  };
  __shared__ RowPtrs s_ptrs_all[NW];
  volatile RowPtrs* const pp = &s_ptrs_all[wi];
  // two-row delay line of forward records: [slot][field][lane], written in stage A of row r and read back by
  // the same lane in stages E/F two steps later (no synchronisation needed).  In registers these 51 values
  // pushed the kernel over its 168-register budget (17-27 local-memory spills per three rows).
  __shared__ float sRec[(GRAD && SFM_SSIM_SREC) ? NW * 3 * 18 * 32 : 1];
  const int lane = threadIdx.x & 31;
  float* const myrec = sRec + wi * (3 * 18 * 32) + lane;
  const StripTask t = decode_strip(p, blockIdx.x);
  const int s = t.s, b = t.b, h = p.h[s], w = p.w[s], S = p.S;
  const Geo geo = make_geo(p, s);
  const int xx = t.x0 - 2 + lane;                       // this lane's image column (may be outside)
  const bool col_in = (xx >= 0) && (xx < w);
  const bool col_own = (lane >= 2) && (lane < 2 + SSIM_IW) && (xx < w);
  const float gyv = p.gy ? __ldg(p.gy) : 1.f;
  const float inv_n3 = p.inv_n3[s];
  const float wpix = gyv * (1.f - p.ssim_rate) * inv_n3;
  const float wssim = gyv * p.ssim_rate * inv_n3;
  // SSIM on raw 3x3 window SUMS (9 x the means): the 1/81 factors of numerator and denominator cancel
  const float C1 = 81.f * (0.01f * 0.01f), C2 = 81.f * (0.03f * 0.03f);
  // everything below reads what the prologue kernel wrote (pyramid, tables, the smoothness tasks' gdisp)
  cudaGridDependencySynchronize();
  cudaTriggerProgrammaticLaunchCompletion();           // lets the epilogue's CTAs become resident while this grid drains (after the
                                                       // wait: the epilogue reads the prologue's smoothness partials ahead of its own wait)
  const float* __restrict__ kinvp = p.kinv + ((size_t)b * p.ns + s) * 9;
  const float xf = (float)xx;
  const bool raw = RAW && ((p.raw_disp_mask >> s) & 1u);    // RAW: some scale takes the pre-activation disparity map
  const float kk0 = __ldg(kinvp + 0), kk1 = __ldg(kinvp + 1), kk2 = __ldg(kinvp + 2);
  const float kk3 = __ldg(kinvp + 3), kk4 = __ldg(kinvp + 4), kk5 = __ldg(kinvp + 5);
  const float kk6 = __ldg(kinvp + 6), kk7 = __ldg(kinvp + 7), kk8 = __ldg(kinvp + 8);
  // ray = Kinv.(x, y, 1): r_k = (k_k0*x + k_k1*y) + k_k2 ; the x products are row invariant
  const float rxx = __fmul_rn(kk0, xf), ryx = __fmul_rn(kk3, xf), rzx = __fmul_rn(kk6, xf);
  const int plane = h * w;
  const float* __restrict__ disp = p.disp[s] + (size_t)b * plane;
  const float* __restrict__ tgt = p.tgt_pl[s] + (size_t)b * 3 * plane;
  float* __restrict__ gdisp = GRAD ? p.gdisp[s] + (size_t)b * plane : nullptr;
  const size_t src_img = (size_t)3 * plane;
  float pix_part = 0.f, ssim_part = 0.f;
  const int r_begin = t.y0 - 2, r_end = t.y1 + 2;      // rows [r_begin, r_end) are warped
  const bool opaque_true = p.hseg != 0x7fffffff;       // always true, unknown to ptxas: basic-block fence (see step)

  for (int i = (NW > 1) ? wi : 0; i < ((NW > 1) ? wi + 1 : S); ++i) {
    __syncwarp();
    if (lane < 12) reinterpret_cast<float*>(sP)[lane] = __ldg(p.proj + (((size_t)b * S + i) * p.ns + s) * 12 + lane);
    __syncwarp();
    const float P3 = sP[0].w, P7 = sP[1].w, P11 = sP[2].w;
    float accA[3] = {0.f, 0.f, 0.f}, accB[3] = {0.f, 0.f, 0.f}, accC[3] = {0.f, 0.f, 0.f};
    const bool first = (i == 0);
    const bool writer = (NW == 1) || (wi == 0);          // the warp that loads / stores gdisp
    if (lane == 0) {
      s_ptrs_all[wi].img = p.src_pl[s] + ((size_t)b * S + i) * src_img;
      s_ptrs_all[wi].tgt = tgt;
      s_ptrs_all[wi].disp = disp;
      s_ptrs_all[wi].gdisp = gdisp;
    }
    __syncwarp();
    // rings: window row sums (P, P^2, P.T, T, T^2 per channel), gradient-field row sums, forward records
    V3 hs[3][5], gs[3][3];        // [ring slot][field]: (P, P^2, P.T, T, T^2) and (g_a, g_s, g_c) row sums, three channels each
    Rec rec[3];
    bool mask[3];                                        // all-zero mask of the row's pixel (base_model.py:96)
    float dq[3];                                         // disparity / target rows fetched ahead
    float Tq[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int q = 0; q < 5; ++q) hs[k][q] = v3s(0.f);
#pragma unroll
      for (int q = 0; q < 3; ++q) gs[k][q] = v3s(0.f);
      mask[k] = true;
#pragma unroll
      for (int c = 0; c < 3; ++c) rec[k].Ix[c] = rec[k].Iy[c] = rec[k].P[c] = rec[k].T[c] = 0.f;
      rec[k].q0 = rec[k].q1 = rec[k].q2 = rec[k].r = rec[k].depth = rec[k].dsc = 0.f;
      if (GRAD && SFM_SSIM_SREC) {
#pragma unroll
        for (int q = 0; q < 18; ++q) myrec[(k * 18 + q) * 32] = 0.f;
      }
      dq[k] = 1.f;
      Tq[k][0] = Tq[k][1] = Tq[k][2] = 0.f;
    }
    auto load_dT = [&](int r, float& dd, float* TT) {
      dd = 1.f;
      TT[0] = TT[1] = TT[2] = 0.f;
      if (col_in && (r >= 0) && (r < h) && (r < r_end)) {
        const unsigned o = (unsigned)(r * w + xx);
        const float* tb = pp->tgt;
        dd = __ldg(pp->disp + o);
        TT[0] = __ldg(tb + o);
        TT[1] = __ldg(tb + (o + geo.plane));
        TT[2] = __ldg(tb + (o + 2u * geo.plane));
      }
    };
    // pending row: projected, gathers in flight, consumed by the next step
    Taps I;
    float pwa = 0.f, pwb = 0.f, pwc = 0.f, pwd = 0.f, pq0 = 0.f, pq1 = 0.f, pq2 = 0.f, pr = 0.f, pdepth = 0.f, pdsc = 0.f;
    float g_old = 0.f;

    // refill: project row r (its disparity was fetched a step earlier), issue its gathers, fetch row r+1's
    // disparity / target and the partial gdisp of the row whose backward runs in the next step
    auto refill = [&](auto SLOT, const int r) {
      constexpr int sl = decltype(SLOT)::value, nx = (sl + 1) % 3;
      const bool in_img = col_in && (r >= 0) && (r < h) && (r < r_end);
      float d = dq[sl], dact = 1.f;
      if (raw) d = sfm_disp_act(d, dact);      // producer-side fusion: pre-activation disparity map (warp-uniform branch)
      pdepth = rcp_newton(d);
      if (RAW) pdsc = raw ? pdepth * dact : pdepth;
      const float yf = (float)r;
      const float rx = __fadd_rn(__fadd_rn(rxx, __fmul_rn(kk1, yf)), kk2);
      const float ry = __fadd_rn(__fadd_rn(ryx, __fmul_rn(kk4, yf)), kk5);
      const float rz = __fadd_rn(__fadd_rn(rzx, __fmul_rn(kk7, yf)), kk8);
      const float X = __fmul_rn(pdepth, rx), Y = __fmul_rn(pdepth, ry), Z = __fmul_rn(pdepth, rz);
      float P[12];
      {
        const float4 a = sP[0], bb = sP[1], c = sP[2];
        P[0] = a.x; P[1] = a.y; P[2] = a.z; P[3] = a.w;
        P[4] = bb.x; P[5] = bb.y; P[6] = bb.z; P[7] = bb.w;
        P[8] = c.x; P[9] = c.y; P[10] = c.z; P[11] = c.w;
      }
      PairFwd f;
      pair_project(P, X, Y, Z, geo, f, in_img);
      pwa = f.wa; pwb = f.wb; pwc = f.wc; pwd = f.wd;
      pq0 = f.q0; pq1 = f.q1; pq2 = f.q2; pr = f.r;
      gather_taps(pp->img, geo, f, I);
      load_dT(r + 1, dq[nx], Tq[nx]);
      if (GRAD && (ACCUM || !first) && writer) {
        const int rf = r - 2;
        g_old = (col_own && rf >= t.y0 && rf < t.y1) ? ld_global(pp->gdisp + (unsigned)(rf * w + xx)) : 0.f;
      }
      if (DEBUG && in_img && col_own && (r >= t.y0) && (r < t.y1) && p.dbg_u0[s]) {
        const size_t o = ((size_t)b * S + i) * plane + (size_t)r * w + xx;
        SfmCoord c;
        sfm_project(P, X, Y, Z, w, h, geo.hw, geo.hh, c);
        p.dbg_u0[s][o] = f.inb ? (int)(f.idx % (unsigned)w) : c.u0;
        p.dbg_v0[s][o] = f.inb ? (int)(f.idx / (unsigned)w) : c.v0;
        p.dbg_inb[s][o] = f.inb ? 1 : 0;
      }
    };

    auto step = [&](auto PH, const int r) {
      constexpr int cur = decltype(PH)::value, pv1 = (cur + 2) % 3, pv2 = (cur + 1) % 3;
      const int rf = r - 2;
      const bool do_f = col_own && rf >= t.y0 && rf < t.y1;
      const float g_prev = g_old;
      // ---------------- stage A: blend pixel (r, lane) from the taps gathered during the previous step
      Rec& rc_ = rec[cur];
      {
        const float* T = Tq[cur];
        const float w1 = __fmul_rn(pwa, pwc), w2 = __fmul_rn(pwb, pwc);
        const float w3 = __fmul_rn(pwa, pwd), w4 = __fmul_rn(pwb, pwd);
        const float P0 = sfm_blend(w1, w2, w3, w4, I.a[0], I.b[0], I.c[0], I.d[0]);
        const float P1 = sfm_blend(w1, w2, w3, w4, I.a[1], I.b[1], I.c[1], I.d[1]);
        const float P2 = sfm_blend(w1, w2, w3, w4, I.a[2], I.b[2], I.c[2], I.d[2]);
        const bool m = (P0 == 0.f) && (P1 == 0.f) && (P2 == 0.f);               // true outside the image / view
        const bool own = col_own && (r >= t.y0) && (r < t.y1);
        pix_part += (own && !m) ? (fabsf(P0 - T[0]) + fabsf(P1 - T[1]) + fabsf(P2 - T[2])) : 0.f;
        mask[cur] = m;
        rc_.P[0] = P0; rc_.P[1] = P1; rc_.P[2] = P2;
        rc_.T[0] = T[0]; rc_.T[1] = T[1]; rc_.T[2] = T[2];
        if (GRAD) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            rc_.Ix[c] = pwc * (I.b[c] - I.a[c]) + pwd * (I.d[c] - I.c[c]);
            rc_.Iy[c] = pwa * (I.c[c] - I.a[c]) + pwb * (I.d[c] - I.b[c]);
          }
          rc_.q0 = pq0; rc_.q1 = pq1; rc_.q2 = pq2; rc_.r = pr;
          rc_.depth = pdepth;
          rc_.dsc = RAW ? pdsc : pdepth;
          if (SFM_SSIM_SREC) {
            float* o = myrec + cur * 18 * 32;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              o[(0 + c) * 32] = rc_.Ix[c];
              o[(3 + c) * 32] = rc_.Iy[c];
              o[(6 + c) * 32] = rc_.P[c];
              o[(9 + c) * 32] = rc_.T[c];
            }
            o[12 * 32] = pq0; o[13 * 32] = pq1; o[14 * 32] = pq2; o[15 * 32] = pr; o[16 * 32] = pdepth;
            if (RAW) o[17 * 32] = pdsc;
          }
        }
        if (DEBUG && own && p.dbg_P[s]) {
          float* o = p.dbg_P[s] + ((size_t)b * S + i) * 3 * plane + (size_t)r * w + xx;
          o[0] = P0;
          o[plane] = P1;
          o[2 * (size_t)plane] = P2;
        }
      }
      // ---------------- refill for row r + 1: its gathers fly while stages B..F of this step run
      refill(IC<pv2>{}, r + 1);          // slot of row r+1 in the 3-ring == (cur + 1) % 3
      // Basic-block fence.  The refill has no consumer before the next step, so ptxas' critical-path scheduler
      // otherwise sinks the projection and the four gathers to the end of the step (seen in the SASS; ncu:
      // half of all stall samples were long-scoreboard waits on the taps ~60 instructions after their issue).
      // A branch it cannot fold keeps them ahead of stages B..F, which then cover the latency.
#if SFM_SSIM_FENCE
      if (opaque_true) {
#endif
      // ---------------- stage B: row sums of P, P^2, P.T, T, T^2 over lanes-1..+1 (zero outside the image)
      {
        const V3 pv = v3(rc_.P[0], rc_.P[1], rc_.P[2]), tv = v3(rc_.T[0], rc_.T[1], rc_.T[2]);
        const V3 pl = vshfl_up(pv), prr = vshfl_down(pv), tl = vshfl_up(tv), tr = vshfl_down(tv);
        V3* h0 = hs[cur];
        h0[0] = (pl + pv) + prr;
        h0[1] = vfma(prr, prr, vfma(pv, pv, pl * pl));
        h0[2] = vfma(prr, tr, vfma(pv, tv, pl * tl));
        h0[3] = (tl + tv) + tr;
        h0[4] = vfma(tr, tr, vfma(tv, tv, tl * tl));
      }
      // ---------------- stage C: SSIM at (rc = r-1, lane) from rows r-2, r-1, r
      const int rc = r - 1;
      V3 g0[3];
      {
        const V3* h0 = hs[cur];
        const V3* h1 = hs[pv1];
        const V3* h2 = hs[pv2];
        const bool c_in = col_in && (rc >= 0) && (rc < h) && (r >= r_begin + 2);
        const bool live_px = c_in && !mask[pv1];
        const bool own_c = col_own && (rc >= t.y0) && (rc < t.y1);
        const float lw = (live_px && own_c) ? 1.f : 0.f;
        // window sums: Sp = 9a, Spp = 9 A(P^2), Spt = 9 A(PT), St = 9 my, Stt = 9 A(T^2)
        const V3 Sp = (h2[0] + h1[0]) + h0[0];
        const V3 Spp = (h2[1] + h1[1]) + h0[1];
        const V3 Spt = (h2[2] + h1[2]) + h0[2];
        const V3 St = (h2[3] + h1[3]) + h0[3];
        const V3 Stt = (h2[4] + h1[4]) + h0[4];
        // 81 x the reference's quantities: n1 = 2 a my + c1, n2 = 2 sxy + c2, d1 = a^2 + my^2 + c1, d2 = sx + sy + c2
        const V3 k2 = v3s(2.f), k9 = v3s(9.f), kC1 = v3s(C1), kC2 = v3s(C2);
        const V3 pp = Sp * Sp, tt = St * St, pt = Sp * St;
        const V3 n1 = vfma(k2, pt, kC1);
        const V3 n2 = vfma(k2, vfma(k9, Spt, vneg(pt)), kC2);
        const V3 d1v = (pp + tt) + kC1;
        const V3 d2v = (vfma(k9, Spp, vneg(pp)) + vfma(k9, Stt, vneg(tt))) + kC2;
        const V3 n = n1 * n2, dd = d1v * d2v;
        const V3 rd = v3(rcp_approx(dd.a.x), rcp_approx(dd.a.y), rcp_approx(dd.b));
        const V3 q = n * rd;                          // SSIM
        const V3 raw = vfma(v3s(-0.5f), q, v3s(0.5f));
        ssim_part = fmaf((__saturatef(raw.a.x) + __saturatef(raw.a.y)) + __saturatef(raw.b), lw, ssim_part);
        if (GRAD) {
          // d raw / d(Sp, Spp, Spt) with n, d in the 81x units (SURVEY A.7 rescaled):
          //   g_n = dL/dn = -0.5 w / d ; g_d = dL/dd = 0.5 w n / d^2
          //   dn/dSp = 2 St (n2 - n1) ; dn/dSpt = 18 n1 ; dd/dSp = 2 Sp (d2 - d1) ; dd/dSpp = 9 d1
          // F.clip passes gradient inside [0, 1]
          const float wl = live_px ? -0.5f * wssim : 0.f;
          const V3 wv = v3((raw.a.x >= 0.f && raw.a.x <= 1.f) ? wl : 0.f, (raw.a.y >= 0.f && raw.a.y <= 1.f) ? wl : 0.f,
                           (raw.b >= 0.f && raw.b <= 1.f) ? wl : 0.f);
          const V3 g_n = wv * rd;
          const V3 g_d = vneg(g_n) * q;
          g0[0] = vfma(g_n * St, n2 + vneg(n1), (g_d * Sp) * (d2v + vneg(d1v)));   // (dL/dSp) / 2
          g0[1] = g_d * d1v;                                                       // (dL/dSpp) / 9
          g0[2] = g_n * n1;                                                        // (dL/dSpt) / 18
        }
      }
      if (GRAD) {
        // ---------------- stage D: row sums of the three gradient fields of row rc
        V3* gh0 = gs[cur];
#pragma unroll
        for (int q = 0; q < 3; ++q) gh0[q] = (vshfl_up(g0[q]) + g0[q]) + vshfl_down(g0[q]);
        // ---------------- stages E + F: dL/dP and the warp backward for pixel (rf = r-2, lane)
        const V3* g1 = gs[pv1];
        const V3* g2 = gs[pv2];
        Rec rb = rec[pv2];
        if (SFM_SSIM_SREC) {
          const float* o = myrec + pv2 * 18 * 32;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            rb.Ix[c] = o[(0 + c) * 32];
            rb.Iy[c] = o[(3 + c) * 32];
            rb.P[c] = o[(6 + c) * 32];
            rb.T[c] = o[(9 + c) * 32];
          }
          rb.q0 = o[12 * 32]; rb.q1 = o[13 * 32]; rb.q2 = o[14 * 32]; rb.r = o[15 * 32]; rb.depth = o[16 * 32];
          rb.dsc = RAW ? o[17 * 32] : rb.depth;
        }
        const bool mf = mask[pv2];
        float gP[3];
        {
          const V3 Aa = (g2[0] + g1[0]) + gh0[0];
          const V3 As = (g2[1] + g1[1]) + gh0[1];
          const V3 Ac = (g2[2] + g1[2]) + gh0[2];
          // dL/dP = sum over the 3x3 windows containing the pixel of dL/dSp + 2P dL/dSpp + T dL/dSpt
          //       = 2 Aa + 18 P As + 18 T Ac
          const V3 Pb = v3(rb.P[0], rb.P[1], rb.P[2]), Tb = v3(rb.T[0], rb.T[1], rb.T[2]);
          const V3 gsv = vfma(v3s(18.f), vfma(Tb, Ac, Pb * As), v3s(2.f) * Aa);
          const float gs3[3] = {gsv.a.x, gsv.a.y, gsv.b};
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float gl1 = mf ? 0.f : sign_times(rb.P[c] - rb.T[c], wpix);
            gP[c] = do_f ? gs3[c] + gl1 : 0.f;
          }
        }
        // sampler + projection backward; out-of-view pixels have Ix = Iy = 0 and r = 0
        const float gu = gP[0] * rb.Ix[0] + gP[1] * rb.Ix[1] + gP[2] * rb.Ix[2];
        const float gv = gP[0] * rb.Iy[0] + gP[1] * rb.Iy[1] + gP[2] * rb.Iy[2];
        const float gq0 = gu * rb.r, gq1 = gv * rb.r;
        const float gq2 = -(gq0 * rb.q0 + gq1 * rb.q1) * rb.r;
        const float gdd = gq0 * (rb.q0 - P3) + gq1 * (rb.q1 - P7) + gq2 * (rb.q2 - P11);
        const float yfb = (float)rf;
        const float e0 = gq0 * rb.depth, e1 = gq1 * rb.depth, e2 = gq2 * rb.depth;
        accA[0] += e0; accA[1] += e1; accA[2] += e2;
        accB[0] = fmaf(e0, yfb, accB[0]); accB[1] = fmaf(e1, yfb, accB[1]); accB[2] = fmaf(e2, yfb, accB[2]);
        accC[0] += gq0; accC[1] += gq1; accC[2] += gq2;
        float* gp = pp->gdisp + (unsigned)(rf * w + xx);
        float gval = g_prev - gdd * rb.dsc;
        if (NW > 1) {
          // warp 1 -> warp 0: (gdd, dsc) of the second source for this row; warp 0 continues the sequential chain
          float* slot = sG + cur * 64 + lane;
          if (wi == 1) { slot[0] = gdd; slot[32] = rb.dsc; }
          __syncthreads();
          if (wi == 0) gval = gval - slot[0] * slot[32];
        }
        if (do_f && writer) st_global(gp, gval);
      }
#if SFM_SSIM_FENCE
      }
#endif
    };

    // prologue: rows r_begin (slot 0) and its lookahead
    load_dT(r_begin, dq[0], Tq[0]);
    refill(IC<0>{}, r_begin);
    // the row loop always runs whole groups of three steps (ring rotation = renaming); rows past r_end are
    // out of range everywhere (loads predicated off, nothing owned)
#pragma unroll 1
    for (int r = r_begin; r < r_end; r += 3) {
      step(IC<0>{}, r);
      step(IC<1>{}, r + 1);
      step(IC<2>{}, r + 2);
    }
    // ---- flush: expand (A, B, C) to dL/dP = sum gq (x) (X, Y, Z, 1) with X = depth*((k0 x + k2) + k1 y) ...
    const bool last = (NW > 1) || (i == S - 1);
    if (GRAD || last) {
      float acc[12];
      if (GRAD) {
        const float cx = rxx + kk2, cy = ryx + kk5, cz = rzx + kk8;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          acc[k * 4 + 0] = cx * accA[k] + kk1 * accB[k];
          acc[k * 4 + 1] = cy * accA[k] + kk4 * accB[k];
          acc[k * 4 + 2] = cz * accA[k] + kk7 * accB[k];
          acc[k * 4 + 3] = accC[k];
        }
      }
      flush_dP(p, acc, b, i, s, lane, last ? pix_part * inv_n3 : 0.f, last ? ssim_part * inv_n3 : 0.f, 0.f, 0, 3, -1, GRAD);
    }
  }
}

}  // namespace

template <typename K>
static int launch_ssim_kernel(K kernel, const SfmFusedParams& p, cudaStream_t stream, int nw = 1) {
  const int n_tasks = p.task_begin[SFM_MAX_SCALES];
  if (sfm_ev_start) SFM_CUDA_CHECK(cudaEventRecord(sfm_ev_start, stream));
  SFM_CUDA_CHECK(sfm_launch_kernel(kernel, n_tasks, 32 * nw, stream, !sfm_ev_start, p));
  if (sfm_ev_stop) SFM_CUDA_CHECK(cudaEventRecord(sfm_ev_stop, stream));
  return 0;
}

int launch_epilogue(const SfmFusedParams& p, cudaStream_t stream);

static int sfm_launch_ssim(SfmFusedParams& p, bool grad, bool accum, bool debug, long long want_warps, cudaStream_t stream) {
  // Task height (rows per strip segment).  Cost model, calibrated on B200 (tools/time_graph.py sweeps):
  // a task marches 3*ceil((hseg+4)/3) rows per source (4 halo rows, row loop unrolled by three); up to
  // SFM_MINB_SSIM single-warp CTAs are resident per SM, i.e. 3 per scheduler.  Above one wave the time is
  // waves x rows; below one wave the busiest scheduler holds w = ceil(tasks / schedulers) warps and runs them at
  // a relative issue efficiency of 0.66 / 0.9 / 1.0 for w = 1 / 2 / 3 (a lone warp cannot hide its own latencies).
  // Example cfg2 (B=4, 128x416): hseg 8 -> 1296 tasks, w=3, 12 rows: 36.8 us; hseg 11 -> 976 tasks, w=2, 15 rows: 33.4 us.
  // With two sources and a sub-wave grid the sources can also be split over the two warps of a CTA (NW = 2): a task
  // is then one pass long and there are twice as many warps, so the same occupancy is reached with taller segments
  // (fewer halo rows); the per-row CTA barrier is charged 5 %.
  (void)want_warps;
  int hseg = 64, nw = 1;
  bool srec = true;       // forward records in shared memory; false: the register-record instances (plain GRAD mode only)
  {
    const long long sched = (long long)g_num_sms * 4, slots = (long long)g_num_sms * SFM_MINB_SSIM;
    const long long slots_reg = (long long)g_num_sms * SFM_MINB_SSIM_REG;
    const bool may_split = grad && !debug && p.raw_disp_mask == 0 && p.S == 2 && !getenv("SFM_SSIM_NOSPLIT");
    const bool may_reg = grad && !debug && p.raw_disp_mask == 0;
    double best = 1e300;
    for (int var = 0; var <= (may_reg ? 1 : 0); ++var)
      for (int cand = 1; cand <= (may_split ? 2 : 1); ++cand)
        for (int h = 8; h <= 64; ++h) {
          long long n = 0;
          for (int s = 0; s < p.ns; ++s) n += (long long)p.B * ((p.w[s] + SSIM_IW - 1) / SSIM_IW) * ((p.h[s] + h - 1) / h);
          const long long warps = n * cand;
          const double rows = 3.0 * ((h + 4 + 2) / 3) * (cand == 1 ? p.S : 1) * (cand == 1 ? 1.0 : 1.05);
          double cost;
          if (var == 1) {
            // register-record instances: 2 warps per scheduler at most, 9 % faster per row; only while the whole grid is
            // resident at once (above that the shared-memory instances' third warp per scheduler wins: measured at cfg5)
            if (warps > slots_reg) continue;
            const long long w = (warps + sched - 1) / sched;
            cost = 0.91 * rows * (double)w / (w <= 1 ? 0.66 : 0.9);
          } else if (warps > slots) {
            cost = rows * 3.0 * (double)((warps + slots - 1) / slots);
          } else {
            const long long w = (warps + sched - 1) / sched;
            cost = rows * (double)w / (w <= 1 ? 0.66 : w == 2 ? 0.9 : 1.0);
          }
          if (cost < best || (cost == best && cand == nw)) { best = cost; hseg = h; nw = cand; srec = (var == 0); }
        }
  }
  if (srec && nw == 1) {
    // Grids of many waves: the time follows the total number of row steps (measured at cfg5: 64 / 86 / 128 rows per
    // segment -> 934 / 933 / 912 us for 276 / 276 / 270 row steps per strip), so a taller segment -- fewer halo rows --
    // is taken when it saves at least 1.5 % of the row steps and still leaves three waves of tasks.
    const long long slots = (long long)g_num_sms * SFM_MINB_SSIM;
    auto work = [&](int h, long long* n_out) {
      long long n = 0, wk = 0;
      for (int s = 0; s < p.ns; ++s) {
        const long long strips = (long long)p.B * ((p.w[s] + SSIM_IW - 1) / SSIM_IW);
        const int full = p.h[s] / h, rem = p.h[s] - full * h;
        n += strips * (full + (rem ? 1 : 0));
        wk += strips * ((long long)full * (3 * ((h + 4 + 2) / 3)) + (rem ? 3 * ((rem + 4 + 2) / 3) : 0));
      }
      *n_out = n;
      return wk;
    };
    long long n0 = 0;
    long long w0 = work(hseg, &n0);
    if (n0 > 3 * slots)
      for (int h2 : {96, 128, 192, 256}) {
        long long n2 = 0;
        const long long w2 = work(h2, &n2);
        if (n2 >= 3 * slots && (double)w2 < 0.985 * (double)w0) { hseg = h2; w0 = w2; }
      }
  }
  {
    const char* e = getenv("SFM_HSEG");          // development knobs
    if (e && atoi(e) > 0) hseg = atoi(e);
    const char* f = getenv("SFM_SSIM_NW");
    if (f && (atoi(f) == 1 || (atoi(f) == 2 && grad && !debug && p.raw_disp_mask == 0 && p.S == 2))) nw = atoi(f);
    const char* g = getenv("SFM_SSIM_SREC");
    if (g && (atoi(g) == 1 || (atoi(g) == 0 && grad && !debug && p.raw_disp_mask == 0))) srec = atoi(g) != 0;
  }
  p.hseg = hseg;
  int total = 0;
  for (int s = 0; s < SFM_MAX_SCALES; ++s) {
    p.task_begin[s] = total;
    if (s < p.ns) {
      p.nstrip[s] = (p.w[s] + SSIM_IW - 1) / SSIM_IW;
      p.nseg[s] = (p.h[s] + hseg - 1) / hseg;
      total += p.B * p.nstrip[s] * p.nseg[s];
    } else {
      p.nstrip[s] = p.nseg[s] = 1;
    }
  }
  p.task_begin[SFM_MAX_SCALES] = total;
  int rc;
  const bool raw = p.raw_disp_mask != 0;
  if (nw == 2) {
    if (srec)
      rc = accum ? launch_ssim_kernel(sfm_ssim_march_kernel<true, true, false, false, 2, true>, p, stream, 2)
                 : launch_ssim_kernel(sfm_ssim_march_kernel<true, false, false, false, 2, true>, p, stream, 2);
    else
      rc = accum ? launch_ssim_kernel(sfm_ssim_march_kernel<true, true, false, false, 2, false>, p, stream, 2)
                 : launch_ssim_kernel(sfm_ssim_march_kernel<true, false, false, false, 2, false>, p, stream, 2);
    if (rc) return rc;
    return launch_epilogue(p, stream);
  }
  if (!srec) {
    rc = accum ? launch_ssim_kernel(sfm_ssim_march_kernel<true, true, false, false, 1, false>, p, stream)
               : launch_ssim_kernel(sfm_ssim_march_kernel<true, false, false, false, 1, false>, p, stream);
    if (rc) return rc;
    return launch_epilogue(p, stream);
  }
#define SFM_SSIM(GR, AC)                                                                                   \
  rc = raw ? (debug ? launch_ssim_kernel(sfm_ssim_march_kernel<GR, AC, true, true, 1>, p, stream)             \
                    : launch_ssim_kernel(sfm_ssim_march_kernel<GR, AC, false, true, 1>, p, stream))           \
           : (debug ? launch_ssim_kernel(sfm_ssim_march_kernel<GR, AC, true, false, 1>, p, stream)            \
                    : launch_ssim_kernel(sfm_ssim_march_kernel<GR, AC, false, false, 1>, p, stream))
  if (grad) {
    if (accum) { SFM_SSIM(true, true); } else { SFM_SSIM(true, false); }
  } else {
    SFM_SSIM(false, false);
  }
#undef SFM_SSIM
  if (rc) return rc;
  return launch_epilogue(p, stream);
}
