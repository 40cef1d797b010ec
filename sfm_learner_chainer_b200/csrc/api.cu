// api.cu -- extern "C" entry points of libsfmloss.so (see include/sfmloss.h for the contract and the
// reference interfaces each one replaces).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void sfm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sfm_last_error(void) { return g_err; }

thread_local cudaEvent_t sfm_ev_start = nullptr, sfm_ev_stop = nullptr;   // recorded around the fused kernel (fused_loss.cu)
extern "C" int sfm_set_kernel_events(void* start_event, void* stop_event) {
  sfm_ev_start = (cudaEvent_t)start_event;
  sfm_ev_stop = (cudaEvent_t)stop_event;
  return 0;
}
extern "C" int sfm_version(void) { return SFM_VERSION; }

bool sfm_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SFM_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

int sfm_validate_desc(const SfmDesc* d) {
  if (!d) { sfm_set_error("SfmDesc is NULL"); return SFM_E_NULL_POINTER; }
  if (d->B < 1 || d->S < 1 || d->S > SFM_MAX_SOURCES || d->n_scales < 1 || d->n_scales > SFM_MAX_SCALES) {
    sfm_set_error("invalid SfmDesc: B=%d S=%d (1..%d) n_scales=%d (1..%d)", d->B, d->S, SFM_MAX_SOURCES, d->n_scales, SFM_MAX_SCALES);
    return SFM_E_INVALID_DESC;
  }
  if (d->B_global != 0 && d->B_global < d->B) {
    sfm_set_error("invalid SfmDesc: B_global=%d < B=%d", d->B_global, d->B);
    return SFM_E_INVALID_DESC;
  }
  const int hs = d->H >> (d->n_scales - 1), ws = d->W >> (d->n_scales - 1);
  if (d->H < 2 || d->W < 2 || hs < 4 || ws < 4) {
    // resize_images needs >= 2 input rows/cols; the smoothness means need (h-2)*(w-2) > 0 at every scale; and
    // with >= 4 rows/cols every tap of an out-of-view pixel (x2 rule) lies in the sampler's zero padding
    sfm_set_error("invalid shape: H=%d W=%d give %dx%d at the coarsest of %d scales (need >= 4x4)", d->H, d->W, hs, ws, d->n_scales);
    return SFM_E_INVALID_SHAPE;
  }
  if ((long long)d->B * (1 + d->S) * d->H * d->W >= (1ll << 31)) {
    sfm_set_error("invalid shape: B*(1+S)*H*W exceeds 2^31 pixels");
    return SFM_E_INVALID_SHAPE;
  }
  if (d->raw_disp_scales >> d->n_scales) {
    sfm_set_error("invalid SfmDesc: raw_disp_scales=0x%x names a scale >= n_scales=%d", d->raw_disp_scales, d->n_scales);
    return SFM_E_INVALID_DESC;
  }
  if (d->raw_pose_hw < 0 || d->raw_pose_hw > 128) {
    sfm_set_error("unsupported SfmDesc: raw_pose_hw=%d (0 = 6-DoF vectors, 1..128 = positions of the poseout map)", d->raw_pose_hw);
    return SFM_E_UNSUPPORTED;
  }
  if (d->raw_pose_hw > 0 && (d->flags & SFM_FLAG_TABLES_PROVIDED)) {
    sfm_set_error("unsupported SfmDesc: raw_pose_hw > 0 together with SFM_FLAG_TABLES_PROVIDED (the tables are built from the reduced poses)");
    return SFM_E_UNSUPPORTED;
  }
  if (!(d->smooth_reg == d->smooth_reg) || !(d->exp_reg == d->exp_reg) || !(d->ssim_rate == d->ssim_rate)) {
    sfm_set_error("invalid SfmDesc: NaN loss weight");
    return SFM_E_INVALID_DESC;
  }
  return 0;
}

extern "C" size_t sfm_workspace_bytes(const SfmDesc* desc) {
  if (sfm_validate_desc(desc) != 0) return 0;
  SfmWsLayout L;
  sfm_ws_layout(desc, &L);
  return L.total;
}

// ------------------------------------------------------------------------------------------------
// common driver
// ------------------------------------------------------------------------------------------------
namespace {

struct Modes {
  bool use_exp, use_ssim, use_smooth;
};

// Python truthiness of the reference's flags (base_model.py:75,86,103,112): 0 / 0.0 disable a term;
// SSIM sits in the else-branch of the explainability test (:103-115).
Modes modes_of(const SfmDesc* d) {
  Modes m;
  m.use_exp = d->exp_reg != 0.f;
  m.use_ssim = !m.use_exp && d->ssim_rate != 0.f;
  m.use_smooth = d->smooth_reg != 0.f;
  return m;
}

int check_inputs(const SfmDesc* d, const SfmInputs* in, bool need_images) {
  if (!in) { sfm_set_error("SfmInputs is NULL"); return SFM_E_NULL_POINTER; }
  const Modes m = modes_of(d);
  if (need_images && (!in->tgt || !in->src)) { sfm_set_error("tgt/src image pointer is NULL"); return SFM_E_NULL_POINTER; }
  if (!in->intrinsics || !in->poses) { sfm_set_error("intrinsics/poses pointer is NULL"); return SFM_E_NULL_POINTER; }
  for (int s = 0; s < d->n_scales; ++s) {
    if (!in->disps[s]) { sfm_set_error("disps[%d] is NULL", s); return SFM_E_NULL_POINTER; }
    if (m.use_exp && !in->logits[s]) { sfm_set_error("exp_reg > 0 but logits[%d] is NULL", s); return SFM_E_NULL_POINTER; }
  }
  if ((d->flags & SFM_FLAG_TABLES_PROVIDED) && (!in->proj || !in->kinv)) {
    sfm_set_error("SFM_FLAG_TABLES_PROVIDED set but proj/kinv is NULL");
    return SFM_E_NULL_POINTER;
  }
  return 0;
}

int check_grads(const SfmDesc* d, const SfmGrads* g) {
  if (!g) { sfm_set_error("SfmGrads is NULL"); return SFM_E_NULL_POINTER; }
  const Modes m = modes_of(d);
  if (!g->gposes) { sfm_set_error("gposes is NULL"); return SFM_E_NULL_POINTER; }
  for (int s = 0; s < d->n_scales; ++s) {
    if (!g->gdisps[s]) { sfm_set_error("gdisps[%d] is NULL", s); return SFM_E_NULL_POINTER; }
    if (m.use_exp && !g->glogits[s]) { sfm_set_error("exp_reg > 0 but glogits[%d] is NULL", s); return SFM_E_NULL_POINTER; }
  }
  return 0;
}

int run_prep(const SfmDesc* d, const SfmInputs* in, void* workspace, bool do_pyramid, const SfmSmoothParams* sm, int sm_mode,
             cudaStream_t st) {
  SfmWsLayout L;
  sfm_ws_layout(d, &L);
  char* ws = (char*)workspace;
  SfmPrepParams p{};
  p.B = d->B; p.S = d->S; p.H = d->H; p.W = d->W; p.ns = d->n_scales;
  p.do_pyramid = do_pyramid ? 1 : 0;
  p.build_tables = (d->flags & SFM_FLAG_TABLES_PROVIDED) ? 0 : 1;
  p.tgt = in->tgt; p.src = in->src; p.intrinsics = in->intrinsics; p.poses = in->poses;
  for (int s = 0; s < d->n_scales; ++s) {
    p.tgt_pyr[s] = (float*)(ws + L.off_tgt[s]);
    p.src_pyr[s] = (float*)(ws + L.off_src[s]);
  }
  p.proj_out = (float*)(ws + L.off_proj);
  p.kinv_out = (float*)(ws + L.off_kinv);
  p.acc = (double*)(ws + L.off_acc);
  p.n_acc = (int)L.acc_doubles;
  p.raw_pose_hw = d->raw_pose_hw;
  p.posevec_out = (float*)(ws + L.off_posevec);
  return sfm_launch_prep(p, sm, sm_mode, st);
}

int run_loss(const SfmDesc* d, const SfmInputs* in, float* losses_out, const SfmGrads* grads, const float* gy,
             const SfmDebug* dbg, void* workspace, cudaStream_t st, const SfmPeerDev* peer = nullptr) {
  int rc = sfm_validate_desc(d);
  if (rc) return rc;
  const bool reuse = (d->flags & SFM_FLAG_REUSE_PYRAMID) != 0;
  rc = check_inputs(d, in, true);      // the images are always needed: scale 0 is read from the caller's tensors
  if (rc) return rc;
  if (grads && (rc = check_grads(d, grads))) return rc;
  if (!workspace) { sfm_set_error("workspace is NULL"); return SFM_E_NULL_POINTER; }
  if (((uintptr_t)workspace & 255) != 0) { sfm_set_error("workspace must be 256-byte aligned"); return SFM_E_INVALID_DESC; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    sfm_set_error("no CUDA device: libsfmloss has no CPU fallback");
    return SFM_E_NO_DEVICE;
  }
  const Modes m = modes_of(d);
  SfmWsLayout L;
  sfm_ws_layout(d, &L);
  char* ws = (char*)workspace;
  SfmFusedParams p{};
  p.B = d->B; p.S = d->S; p.ns = d->n_scales;
  const double Bg = (double)(d->B_global > 0 ? d->B_global : d->B);
  for (int s = 0; s < d->n_scales; ++s) {
    const int h = d->H >> s, w = d->W >> s;
    p.h[s] = h; p.w[s] = w;
    p.tgt_pl[s] = s == 0 ? in->tgt : (const float*)(ws + L.off_tgt[s]);
    p.src_pl[s] = s == 0 ? in->src : (const float*)(ws + L.off_src[s]);
    p.disp[s] = in->disps[s];
    p.logits[s] = m.use_exp ? in->logits[s] : nullptr;
    p.gdisp[s] = grads ? grads->gdisps[s] : nullptr;
    p.glogits[s] = (grads && m.use_exp) ? grads->glogits[s] : nullptr;
    p.inv_n3[s] = (float)(1.0 / (Bg * 3.0 * h * w));
    p.inv_n1[s] = (float)(1.0 / (Bg * h * w));
    const double wgt = (double)d->smooth_reg / (double)(1 << s);
    p.sm_dx2[s] = (float)(wgt / (Bg * h * (w - 2)));
    p.sm_mix[s] = (float)(wgt / (Bg * (h - 1) * (w - 1)));
    p.sm_dy2[s] = (float)(wgt / (Bg * (h - 2) * w));
    p.sm_ex[s] = (float)(wgt / (Bg * h * (w - 1)));
    p.sm_ey[s] = (float)(wgt / (Bg * (h - 1) * w));
    if (dbg) {
      p.dbg_P[s] = dbg->P[s]; p.dbg_u0[s] = dbg->u0[s]; p.dbg_v0[s] = dbg->v0[s]; p.dbg_inb[s] = dbg->inb[s];
    }
  }
  const bool tables = (d->flags & SFM_FLAG_TABLES_PROVIDED) != 0;
  p.proj = tables ? in->proj : (const float*)(ws + L.off_proj);
  p.kinv = tables ? in->kinv : (const float*)(ws + L.off_kinv);
  p.intrinsics = in->intrinsics;
  p.poses = d->raw_pose_hw > 0 ? (const float*)(ws + L.off_posevec) : in->poses;
  p.raw_pose_hw = d->raw_pose_hw;
  p.raw_disp_mask = d->raw_disp_scales;
  p.gy = gy;
  p.acc = (double*)(ws + L.off_acc);
  p.losses_out = losses_out;
  p.peer = SfmPeerDev{};
  if (peer) p.peer = *peer;
  p.gposes = grads ? grads->gposes : nullptr;
  p.smooth_reg = d->smooth_reg;
  p.exp_reg = m.use_exp ? d->exp_reg : 0.f;
  p.ssim_rate = d->ssim_rate;     // enters `total` even when the SSIM term itself is skipped (base_model.py:117)
  p.use_smooth = m.use_smooth ? 1 : 0;
  p.edge_smooth = (d->flags & SFM_FLAG_EDGE_AWARE_SMOOTH) ? 1 : 0;
  int mode = 0;
  if (m.use_exp) mode |= SFM_MODE_EXP;
  if (m.use_ssim) mode |= SFM_MODE_SSIM;
  if (grads) mode |= SFM_MODE_GRAD;
  if (dbg) mode |= SFM_MODE_DEBUG;
  // prologue kernel: pyramid + tables + cell reset and, in the same grid, the second-order smoothness tasks (the
  // edge-aware variant reads the pyramid and is launched by sfm_launch_fused after it)
  SfmSmoothParams sm{};
  int sm_mode = 0;
  p.sm_part = nullptr;
  p.n_sm_part = 0;
  if (m.use_smooth && !p.edge_smooth) {
    sfm_plan_smooth(p, sm);
    sm.part = (float*)(ws + L.off_smpart);
    sm_mode = grads ? 2 : 1;
    p.sm_part = sm.part;
    p.n_sm_part = sm.n_ctas;
  }
  rc = run_prep(d, in, workspace, !reuse, &sm, sm_mode, st);
  if (rc) return rc;
  return sfm_launch_fused(p, mode, st);
}

}  // namespace

extern "C" int sfm_loss_forward(const SfmDesc* desc, const SfmInputs* in, float* losses_out, const SfmDebug* debug,
                                void* workspace, void* stream) {
  if (!losses_out) { sfm_set_error("losses_out is NULL"); return SFM_E_NULL_POINTER; }
  return run_loss(desc, in, losses_out, nullptr, nullptr, debug, workspace, (cudaStream_t)stream);
}

extern "C" int sfm_loss_backward(const SfmDesc* desc, const SfmInputs* in, const float* gy, const SfmGrads* grads,
                                 void* workspace, void* stream) {
  if (!grads) { sfm_set_error("grads is NULL"); return SFM_E_NULL_POINTER; }
  return run_loss(desc, in, nullptr, grads, gy, nullptr, workspace, (cudaStream_t)stream);
}

extern "C" int sfm_loss_forward_backward(const SfmDesc* desc, const SfmInputs* in, float* losses_out,
                                         const SfmGrads* grads, void* workspace, void* stream) {
  if (!losses_out) { sfm_set_error("losses_out is NULL"); return SFM_E_NULL_POINTER; }
  if (!grads) { sfm_set_error("grads is NULL"); return SFM_E_NULL_POINTER; }
  return run_loss(desc, in, losses_out, grads, nullptr, nullptr, workspace, (cudaStream_t)stream);
}

extern "C" int sfm_loss_forward_backward_peer(const SfmDesc* desc, const SfmInputs* in, float* losses_out,
                                              const SfmGrads* grads, void* workspace, SfmPeer* peer, void* stream) {
  if (!losses_out) { sfm_set_error("losses_out is NULL"); return SFM_E_NULL_POINTER; }
  if (!grads) { sfm_set_error("grads is NULL"); return SFM_E_NULL_POINTER; }
  if (!peer) { sfm_set_error("peer is NULL"); return SFM_E_NULL_POINTER; }
  const SfmPeerDev* dev = sfm_peer_dev(peer);
  if (!dev) { sfm_set_error("sfm_loss_forward_backward_peer: the peer object is not connected (sfm_peer_connect)"); return SFM_E_UNSUPPORTED; }
  return run_loss(desc, in, losses_out, grads, nullptr, nullptr, workspace, (cudaStream_t)stream, dev);
}

extern "C" int sfm_scale_grads(const SfmDesc* desc, const float* gy, const SfmGrads* grads, void* stream) {
  int rc = sfm_validate_desc(desc);
  if (rc) return rc;
  if (!gy) { sfm_set_error("gy is NULL"); return SFM_E_NULL_POINTER; }
  if ((rc = check_grads(desc, grads))) return rc;
  const Modes m = modes_of(desc);
  float* ptrs[2 * SFM_MAX_SCALES + 1];
  long long counts[2 * SFM_MAX_SCALES + 1];
  int n = 0;
  for (int s = 0; s < desc->n_scales; ++s) {
    const long long hw = (long long)(desc->H >> s) * (desc->W >> s);
    ptrs[n] = grads->gdisps[s]; counts[n++] = desc->B * hw;
    if (m.use_exp) { ptrs[n] = grads->glogits[s]; counts[n++] = (long long)desc->B * desc->S * hw; }
  }
  ptrs[n] = grads->gposes; counts[n++] = (long long)desc->B * desc->S * 6 * (desc->raw_pose_hw > 0 ? desc->raw_pose_hw : 1);
  return sfm_launch_scale(ptrs, counts, n, gy, (cudaStream_t)stream);
}

extern "C" int sfm_pyramid(const SfmDesc* desc, const float* tgt, const float* src, void* workspace, void* stream) {
  int rc = sfm_validate_desc(desc);
  if (rc) return rc;
  if (!tgt || !src || !workspace) { sfm_set_error("sfm_pyramid: null pointer"); return SFM_E_NULL_POINTER; }
  SfmWsLayout L;
  sfm_ws_layout(desc, &L);
  char* ws = (char*)workspace;
  SfmPrepParams p{};
  p.B = desc->B; p.S = desc->S; p.H = desc->H; p.W = desc->W; p.ns = desc->n_scales;
  p.do_pyramid = 1; p.build_tables = 0;
  p.tgt = tgt; p.src = src;
  for (int s = 0; s < desc->n_scales; ++s) {
    p.tgt_pyr[s] = (float*)(ws + L.off_tgt[s]);
    p.src_pyr[s] = (float*)(ws + L.off_src[s]);
  }
  p.acc = (double*)(ws + L.off_acc);
  p.n_acc = 0;
  return sfm_launch_prep(p, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int sfm_pyramid_export(const SfmDesc* desc, const void* workspace, int scale, float* tgt_out,
                                  float* src_out, void* stream) {
  int rc = sfm_validate_desc(desc);
  if (rc) return rc;
  if (scale < 1 || scale >= desc->n_scales) {
    sfm_set_error("sfm_pyramid_export: scale %d out of range (the workspace holds scales 1..%d; scale 0 is the input itself)", scale,
                  desc->n_scales - 1);
    return SFM_E_INVALID_DESC;
  }
  if (!workspace) { sfm_set_error("sfm_pyramid_export: null workspace"); return SFM_E_NULL_POINTER; }
  SfmWsLayout L;
  sfm_ws_layout(desc, &L);
  const char* ws = (const char*)workspace;
  // the pyramid levels are stored exactly in the output layout: (B,3,h,w) and (B,S,3,h,w)
  const size_t lvl = (size_t)3 * (desc->H >> scale) * (desc->W >> scale) * sizeof(float);
  if (tgt_out) SFM_CUDA_CHECK(cudaMemcpyAsync(tgt_out, ws + L.off_tgt[scale], desc->B * lvl, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (src_out) SFM_CUDA_CHECK(cudaMemcpyAsync(src_out, ws + L.off_src[scale], (size_t)desc->B * desc->S * lvl, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

extern "C" int sfm_build_tables(const SfmDesc* desc, const float* poses, const float* intrinsics, float* proj_out,
                                float* kinv_out, void* stream) {
  int rc = sfm_validate_desc(desc);
  if (rc) return rc;
  if (!poses || !intrinsics || !proj_out || !kinv_out) { sfm_set_error("sfm_build_tables: null pointer"); return SFM_E_NULL_POINTER; }
  SfmPrepParams p{};
  p.B = desc->B; p.S = desc->S; p.H = desc->H; p.W = desc->W; p.ns = desc->n_scales;
  p.do_pyramid = 0; p.build_tables = 1;
  p.intrinsics = intrinsics; p.poses = poses;
  p.proj_out = proj_out; p.kinv_out = kinv_out;
  p.acc = nullptr; p.n_acc = 0;
  p.raw_pose_hw = desc->raw_pose_hw; p.posevec_out = nullptr;
  return sfm_launch_prep(p, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int sfm_ingest_u8(int B, int S, int H, int W, int n_scales, const uint8_t* frames, const float* K_in,
                             const SfmAugment* aug, float* tgt_out, float* src_out, float* intrinsics_out, void* stream) {
  if (B < 1 || S < 1 || S > SFM_MAX_SOURCES || n_scales < 1 || n_scales > SFM_MAX_SCALES) {
    sfm_set_error("sfm_ingest_u8: invalid B=%d S=%d n_scales=%d", B, S, n_scales);
    return SFM_E_INVALID_DESC;
  }
  if (H < 2 || W < 2 || (long long)B * (1 + S) * H * W >= (1ll << 31)) {
    sfm_set_error("sfm_ingest_u8: invalid shape H=%d W=%d", H, W);
    return SFM_E_INVALID_SHAPE;
  }
  if (!frames || !tgt_out || !src_out || (intrinsics_out && !K_in)) { sfm_set_error("sfm_ingest_u8: null pointer"); return SFM_E_NULL_POINTER; }
  return sfm_launch_ingest_u8(B, S, H, W, n_scales, frames, K_in, aug, tgt_out, src_out, intrinsics_out, (cudaStream_t)stream);
}

extern "C" size_t sfm_eval_depth_scratch_bytes(int B, int Hg, int Wg) {
  if (B < 1 || Hg < 1 || Wg < 1) return 0;
  return sfm_eval_scratch_bytes_impl(B, Hg, Wg);
}

extern "C" int sfm_eval_depth(int B, int h, int w, int Hg, int Wg, const float* pred_depth, const float* gt_depth,
                              const uint8_t* mask, float min_depth, float max_depth, float* errors_out, void* scratch,
                              void* stream) {
  if (B < 1 || h < 2 || w < 2 || Hg < 1 || Wg < 1) { sfm_set_error("sfm_eval_depth: invalid shape B=%d %dx%d -> %dx%d", B, h, w, Hg, Wg); return SFM_E_INVALID_SHAPE; }
  if (!(min_depth > 0.f) || !(max_depth >= min_depth)) { sfm_set_error("sfm_eval_depth: need 0 < min_depth <= max_depth"); return SFM_E_INVALID_DESC; }
  if (!pred_depth || !gt_depth || !mask || !errors_out || !scratch) { sfm_set_error("sfm_eval_depth: null pointer"); return SFM_E_NULL_POINTER; }
  return sfm_launch_eval_depth(B, h, w, Hg, Wg, pred_depth, gt_depth, mask, min_depth, max_depth, errors_out, scratch, (cudaStream_t)stream);
}

extern "C" int sfm_disp_activation(long long n, const float* x, float* disp, float* dact, void* stream) {
  if (n < 0) { sfm_set_error("sfm_disp_activation: n < 0"); return SFM_E_INVALID_SHAPE; }
  if (!x || (!disp && !dact)) { sfm_set_error("sfm_disp_activation: null pointer"); return SFM_E_NULL_POINTER; }
  return sfm_launch_disp_activation(n, x, disp, dact, (cudaStream_t)stream);
}

extern "C" int sfm_pose_reduce(int B, int S, int hw, const float* x, float* poses_out, void* stream) {
  if (B < 1 || S < 1 || S > SFM_MAX_SOURCES) { sfm_set_error("sfm_pose_reduce: invalid B=%d S=%d", B, S); return SFM_E_INVALID_DESC; }
  if (hw < 1 || hw > 128) { sfm_set_error("sfm_pose_reduce: unsupported hw=%d (1..128)", hw); return SFM_E_UNSUPPORTED; }
  if (!x || !poses_out) { sfm_set_error("sfm_pose_reduce: null pointer"); return SFM_E_NULL_POINTER; }
  return sfm_launch_pose_reduce(B, S, hw, x, poses_out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// host-buffer path
// ------------------------------------------------------------------------------------------------
struct SfmHostCtx {
  SfmDesc desc;
  cudaStream_t stream;
  void* workspace;
  float *d_tgt, *d_src, *d_K, *d_poses, *d_losses, *d_gposes;
  float* d_disp[SFM_MAX_SCALES];
  float* d_logits[SFM_MAX_SCALES];
  float* d_gdisp[SFM_MAX_SCALES];
  float* d_glogits[SFM_MAX_SCALES];
  // uint8 entry (sfm_loss_step_host_u8_submit): allocated on first use
  uint8_t* d_frames;
  SfmAugment* d_aug;
  float* d_Kin;
  std::vector<void*> allocs;
};

static int host_alloc(SfmHostCtx* c, void** p, size_t bytes) {
  SFM_CUDA_CHECK(cudaMalloc(p, bytes));
  c->allocs.push_back(*p);
  return 0;
}

extern "C" int sfm_host_ctx_destroy(SfmHostCtx* ctx) {
  if (!ctx) return 0;
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" int sfm_host_ctx_create(const SfmDesc* desc, SfmHostCtx** ctx_out) {
  int rc = sfm_validate_desc(desc);
  if (rc) return rc;
  if (!ctx_out) { sfm_set_error("ctx_out is NULL"); return SFM_E_NULL_POINTER; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    sfm_set_error("no CUDA device: libsfmloss has no CPU fallback");
    return SFM_E_NO_DEVICE;
  }
  SfmHostCtx* c = new (std::nothrow) SfmHostCtx();
  if (!c) { sfm_set_error("out of host memory"); return SFM_E_INVALID_DESC; }
  c->desc = *desc;
  c->desc.flags &= ~(SFM_FLAG_TABLES_PROVIDED | SFM_FLAG_REUSE_PYRAMID);
  const SfmDesc* d = &c->desc;
  const Modes m = modes_of(d);
  const size_t img = (size_t)d->H * d->W * 3 * sizeof(float);
#define HA(ptr, bytes)                                                     \
  if ((rc = host_alloc(c, (void**)&(ptr), (bytes)))) { sfm_host_ctx_destroy(c); return rc; }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    sfm_set_error("cudaStreamCreate failed");
    delete c;
    return (int)cudaErrorUnknown;
  }
  HA(c->workspace, sfm_workspace_bytes(d));
  HA(c->d_tgt, d->B * img);
  HA(c->d_src, (size_t)d->B * d->S * img);
  HA(c->d_K, (size_t)d->B * d->n_scales * 9 * sizeof(float));
  const size_t pose_bytes = (size_t)d->B * d->S * 6 * (d->raw_pose_hw > 0 ? d->raw_pose_hw : 1) * sizeof(float);
  HA(c->d_poses, pose_bytes);
  HA(c->d_gposes, pose_bytes);
  HA(c->d_losses, 8 * sizeof(float));
  {
    // the per-scale arrays of one kind are carved back to back out of ONE allocation: when the caller's host arrays
    // are laid out the same way (scale s+1 right behind scale s) they travel in a single copy (host_submit)
    size_t tot = 0;
    for (int s = 0; s < d->n_scales; ++s) tot += (size_t)(d->H >> s) * (d->W >> s);
    float *bd = nullptr, *bg = nullptr, *bl = nullptr, *bgl = nullptr;
    HA(bd, d->B * tot * sizeof(float));
    HA(bg, d->B * tot * sizeof(float));
    if (m.use_exp) {
      HA(bl, (size_t)d->B * d->S * tot * sizeof(float));
      HA(bgl, (size_t)d->B * d->S * tot * sizeof(float));
    }
    size_t off = 0;
    for (int s = 0; s < d->n_scales; ++s) {
      const size_t hw = (size_t)(d->H >> s) * (d->W >> s);
      c->d_disp[s] = bd + d->B * off;
      c->d_gdisp[s] = bg + d->B * off;
      if (m.use_exp) {
        c->d_logits[s] = bl + (size_t)d->B * d->S * off;
        c->d_glogits[s] = bgl + (size_t)d->B * d->S * off;
      }
      off += hw;
    }
  }
#undef HA
  *ctx_out = c;
  return 0;
}

extern "C" int sfm_loss_step_host_wait(SfmHostCtx* c) {
  if (!c) { sfm_set_error("sfm_loss_step_host_wait: null context"); return SFM_E_NULL_POINTER; }
  SFM_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int sfm_loss_step_host(SfmHostCtx* c, const SfmInputs* in, float* losses_out, const SfmGrads* grads) {
  const int rc = sfm_loss_step_host_submit(c, in, losses_out, grads);
  return rc ? rc : sfm_loss_step_host_wait(c);
}

static int host_submit(SfmHostCtx* c, const SfmInputs* in, float* losses_out, const SfmGrads* grads, bool images_on_device);

extern "C" int sfm_loss_step_host_submit(SfmHostCtx* c, const SfmInputs* in, float* losses_out, const SfmGrads* grads) {
  return host_submit(c, in, losses_out, grads, false);
}

extern "C" int sfm_loss_step_host_u8_submit(SfmHostCtx* c, const uint8_t* frames, const float* K_in, const SfmAugment* aug,
                                            const SfmInputs* in, float* losses_out, const SfmGrads* grads) {
  if (!c || !frames || !K_in || !in) { sfm_set_error("sfm_loss_step_host_u8_submit: null pointer"); return SFM_E_NULL_POINTER; }
  const SfmDesc* d = &c->desc;
  const size_t n_frames = (size_t)d->B * (1 + d->S) * d->H * d->W * 3;
  if (aug) {
    // the crop window must lie inside the rescaled image (kitti_raw_transformed.py:36-50); the kernel trusts these
    for (int b = 0; b < d->B; ++b) {
      const SfmAugment& a = aug[b];
      if (a.out_h < d->H || a.out_w < d->W || a.off_y < 0 || a.off_x < 0 || a.off_y > a.out_h - d->H || a.off_x > a.out_w - d->W ||
          !(a.x_scaling > 0.0) || !(a.y_scaling > 0.0)) {
        sfm_set_error("sfm_loss_step_host_u8_submit: aug[%d] is invalid (out %dx%d, offset %d,%d for a %dx%d crop)", b, a.out_h, a.out_w,
                      a.off_y, a.off_x, d->H, d->W);
        return SFM_E_INVALID_DESC;
      }
    }
  }
  // each buffer is guarded on its own: a failed allocation leaves the others usable and is retried by the next call
  int rc0;
  if (!c->d_frames && (rc0 = host_alloc(c, (void**)&c->d_frames, n_frames))) return rc0;
  if (!c->d_aug && (rc0 = host_alloc(c, (void**)&c->d_aug, (size_t)d->B * sizeof(SfmAugment)))) return rc0;
  if (!c->d_Kin && (rc0 = host_alloc(c, (void**)&c->d_Kin, (size_t)d->B * 9 * sizeof(float)))) return rc0;
  cudaStream_t st = c->stream;
  SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_frames, frames, n_frames, cudaMemcpyHostToDevice, st));
  SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_Kin, K_in, (size_t)d->B * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
  if (aug) SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_aug, aug, (size_t)d->B * sizeof(SfmAugment), cudaMemcpyHostToDevice, st));
  int rc = sfm_launch_ingest_u8(d->B, d->S, d->H, d->W, d->n_scales, c->d_frames, c->d_Kin, aug ? c->d_aug : nullptr, c->d_tgt,
                                c->d_src, c->d_K, st);
  if (rc) return rc;
  return host_submit(c, in, losses_out, grads, true);
}

static int host_submit(SfmHostCtx* c, const SfmInputs* in, float* losses_out, const SfmGrads* grads, bool images_on_device) {
  if (!c || !in || !losses_out || !grads) { sfm_set_error("sfm_loss_step_host: null pointer"); return SFM_E_NULL_POINTER; }
  const SfmDesc* d = &c->desc;
  SfmInputs chk = *in;
  if (images_on_device) { chk.tgt = c->d_tgt; chk.src = c->d_src; chk.intrinsics = c->d_K; }   // produced by the ingest kernel
  int rc = check_inputs(d, &chk, true);
  if (rc) return rc;
  if ((rc = check_grads(d, grads))) return rc;
  const Modes m = modes_of(d);
  cudaStream_t st = c->stream;
  const size_t img = (size_t)d->H * d->W * 3 * sizeof(float);
  if (!images_on_device) {
    SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_tgt, in->tgt, d->B * img, cudaMemcpyHostToDevice, st));
    SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_src, in->src, (size_t)d->B * d->S * img, cudaMemcpyHostToDevice, st));
    SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_K, in->intrinsics, (size_t)d->B * d->n_scales * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  const size_t pose_bytes = (size_t)d->B * d->S * 6 * (d->raw_pose_hw > 0 ? d->raw_pose_hw : 1) * sizeof(float);
  SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_poses, in->poses, pose_bytes, cudaMemcpyHostToDevice, st));
  SfmInputs din{};
  SfmGrads dg{};
  din.tgt = c->d_tgt; din.src = c->d_src; din.intrinsics = c->d_K; din.poses = c->d_poses;
  dg.gposes = c->d_gposes;
  // Every copy has a fixed cost of several microseconds (measured on B200: a 1 MB pinned H2D copy reaches 21 GB/s, a
  // 3 MB one 49 GB/s), so host arrays that lie back to back in scale order -- the device side always does -- travel
  // as ONE copy per kind instead of one per scale.
  size_t pix_tot = 0;
  for (int s = 0; s < d->n_scales; ++s) pix_tot += (size_t)(d->H >> s) * (d->W >> s);
  auto contiguous = [&](const float* const* a, size_t per_pixel) {
    for (int s = 0; s + 1 < d->n_scales; ++s)
      if (a[s + 1] != a[s] + per_pixel * (size_t)(d->H >> s) * (d->W >> s)) return false;
    return true;
  };
  const bool disp_1 = contiguous(in->disps, d->B), gdisp_1 = contiguous(grads->gdisps, d->B);
  const bool lg_1 = m.use_exp && contiguous(in->logits, (size_t)d->B * d->S);
  const bool glg_1 = m.use_exp && contiguous(grads->glogits, (size_t)d->B * d->S);
  if (disp_1) SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_disp[0], in->disps[0], d->B * pix_tot * sizeof(float), cudaMemcpyHostToDevice, st));
  if (lg_1) SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_logits[0], in->logits[0], (size_t)d->B * d->S * pix_tot * sizeof(float), cudaMemcpyHostToDevice, st));
  for (int s = 0; s < d->n_scales; ++s) {
    const size_t hw = (size_t)(d->H >> s) * (d->W >> s) * sizeof(float);
    if (!disp_1) SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_disp[s], in->disps[s], d->B * hw, cudaMemcpyHostToDevice, st));
    din.disps[s] = c->d_disp[s];
    dg.gdisps[s] = c->d_gdisp[s];
    if (m.use_exp) {
      if (!lg_1) SFM_CUDA_CHECK(cudaMemcpyAsync(c->d_logits[s], in->logits[s], (size_t)d->B * d->S * hw, cudaMemcpyHostToDevice, st));
      din.logits[s] = c->d_logits[s];
      dg.glogits[s] = c->d_glogits[s];
    }
  }
  rc = sfm_loss_forward_backward(d, &din, c->d_losses, &dg, c->workspace, st);
  if (rc) return rc;
  SFM_CUDA_CHECK(cudaMemcpyAsync(losses_out, c->d_losses, 5 * sizeof(float), cudaMemcpyDeviceToHost, st));
  SFM_CUDA_CHECK(cudaMemcpyAsync(grads->gposes, c->d_gposes, pose_bytes, cudaMemcpyDeviceToHost, st));
  if (gdisp_1) SFM_CUDA_CHECK(cudaMemcpyAsync(grads->gdisps[0], c->d_gdisp[0], d->B * pix_tot * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (glg_1) SFM_CUDA_CHECK(cudaMemcpyAsync(grads->glogits[0], c->d_glogits[0], (size_t)d->B * d->S * pix_tot * sizeof(float), cudaMemcpyDeviceToHost, st));
  for (int s = 0; s < d->n_scales; ++s) {
    const size_t hw = (size_t)(d->H >> s) * (d->W >> s) * sizeof(float);
    if (!gdisp_1) SFM_CUDA_CHECK(cudaMemcpyAsync(grads->gdisps[s], c->d_gdisp[s], d->B * hw, cudaMemcpyDeviceToHost, st));
    if (m.use_exp && !glg_1)
      SFM_CUDA_CHECK(cudaMemcpyAsync(grads->glogits[s], c->d_glogits[s], (size_t)d->B * d->S * hw, cudaMemcpyDeviceToHost, st));
  }
  return 0;
}
