/*
 * sfmloss.h -- C ABI of libsfmloss.so: the SfM-Learner view-synthesis loss path on B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of pfnet/sfm-learner-chainer.  Every entry point
 * cites the reference interface it replaces (paths relative to the reference repo root).
 * The reference-side binding (ctypes + Chainer FunctionNode) is shown in INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / cupy types.
 *  - All tensors are float32, C-contiguous, resident on the CURRENT CUDA device; the caller owns
 *    every buffer (CuPy memory pool, torch allocator, cudaMalloc ...).  The library never frees or
 *    retains caller pointers after a call returns, except inside an SfmHostCtx it created itself.
 *  - Every call is asynchronous on `stream` (0 = legacy default stream), performs no host
 *    synchronisation and no D2H copy, and is capturable in a CUDA graph (the *_host entry point is
 *    the exception: it copies and synchronises by contract).
 *  - Return value: 0 = OK, < 0 = SFM_E_* argument error, > 0 = cudaError_t.  The message of the
 *    last failure on the calling thread is available from sfm_last_error().
 *  - There is NO CPU fallback: without a CUDA device the compute entry points return an error.
 */
#ifndef SFMLOSS_H_
#define SFMLOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_VERSION 102           /* major*100 + minor */
#define SFM_MAX_SCALES 4          /* base_model.py:66 (len(pred_depthes) == 4) */
#define SFM_MAX_SOURCES 8         /* seq_len-1; shipped configs use 2 and 4 */

/* error codes */
#define SFM_OK 0
#define SFM_E_INVALID_DESC (-1)
#define SFM_E_INVALID_SHAPE (-2)
#define SFM_E_NULL_POINTER (-3)
#define SFM_E_UNSUPPORTED (-4)
#define SFM_E_NO_DEVICE (-5)
#define SFM_E_COMM (-6)           /* a collective failed (NCCL error in sfm_last_error) */

/* SfmDesc.flags */
#define SFM_FLAG_TABLES_PROVIDED 0x1u /* caller supplies proj/kinv tables (bit-exact tests; the reference
                                         itself builds P on the host, transform.py:76-90) */
#define SFM_FLAG_REUSE_PYRAMID 0x2u   /* workspace already holds this batch's image pyramid (sfm_pyramid); in->tgt / in->src
                                         are still read (scale 0 is never copied) */
#define SFM_FLAG_NO_TMA 0x4u          /* reserved (accepted and ignored: the marching kernels stage nothing through TMA) */
#define SFM_FLAG_EDGE_AWARE_SMOOTH 0x8u /* the smoothness term is compute_disp_smooth(curr_tgt_img, pred_disps[ns])
                                         (base_model.py:144-155), the edge-aware alternative the reference keeps
                                         commented out at its call site (:78-80), instead of compute_smooth_loss */

/*
 * Problem descriptor.  Mirrors the reference's model flags (base_model.py:37-39:
 * smooth_reg, exp_reg, ssim_rate; experiments/sfm_learner_v1*.yml `architecture:` blocks) and the
 * tensor shapes SFMLearner.__call__ receives (base_model.py:48-58).
 * A flag equal to 0 disables its term (any non-zero value, negative included, enables it) exactly like the
 * Python truthiness tests at base_model.py:75,86,103,112;
 * the explainability branch and SSIM are mutually exclusive as in the reference (:103-115).
 */
typedef struct SfmDesc {
  int32_t B;         /* snippets held by this process                                  */
  int32_t S;         /* source views per snippet (seq_len - 1)                         */
  int32_t H, W;      /* full resolution; scale s is (H >> s, W >> s)                   */
  int32_t n_scales;  /* 1..SFM_MAX_SCALES                                              */
  int32_t B_global;  /* batch every F.mean divides by (0 = B); > B when snippet-sharded */
  float smooth_reg;
  float exp_reg;
  float ssim_rate;
  uint32_t flags;
  /* Producer-side fusion at the seam with the two CNNs (both default to 0 = off):
   * raw_disp_scales  bit s set: disps[s] holds the PRE-activation output of DispNet's `dispout` convolution
   *                  and the kernels apply disp = DISP_SCALING * sigmoid(x) + MIN_DISP themselves
   *                  (models/disp_net.py:7-8,104,110,116,122); gdisps[s] is then the gradient w.r.t. x.
   *                  Scale 0 (disp1) is consumed by the loss alone; disp2..4 are also fed back into the
   *                  decoder (disp_net.py:105,111,117), so a drop-in normally sets bit 0 only.
   * raw_pose_hw      n > 0: poses holds PoseNet's `poseout` map (B, 6*S, n) with n = h'*w' spatial positions
   *                  and the kernels apply pose = 0.01 * mean over n (models/pose_net.py:51-53); gposes then
   *                  has the same (B, 6*S, n) shape.  n <= 128.                                             */
  uint32_t raw_disp_scales;
  int32_t raw_pose_hw;
} SfmDesc;

/* Inputs of one loss evaluation.  Shapes as produced by the reference's nets and dataset:
 *   tgt        (B,3,H,W)            base_model.py:48
 *   src        (B,S,3,H,W)          base_model.py:57-58 (== stacked (B,3S,H,W))
 *   intrinsics (B,n_scales,3,3)     datasets/kitti/kitti_raw_transformed.py:76-93
 *   disps[s]   (B,1,H>>s,W>>s)      models/disp_net.py:124   (disparity, NOT depth; pre-activation map when
 *                                   bit s of desc->raw_disp_scales is set)
 *   poses      (B,S,6)              models/pose_net.py:52-54 (rx,ry,rz,tx,ty,tz per source); (B,6*S,n) raw
 *                                   `poseout` map when desc->raw_pose_hw = n > 0
 *   logits[s]  (B,S,H>>s,W>>s)      models/pose_net.py:56-67; may be NULL when exp_reg == 0
 *   proj       (B,S,n_scales,3,4)   only with SFM_FLAG_TABLES_PROVIDED: K4.T rows 0..2 (transform.py:86-88)
 *   kinv       (B,n_scales,3,3)     only with SFM_FLAG_TABLES_PROVIDED: inverse intrinsics (transform.py:105)
 */
typedef struct SfmInputs {
  const float* tgt;
  const float* src;
  const float* intrinsics;
  const float* disps[SFM_MAX_SCALES];
  const float* poses;
  const float* logits[SFM_MAX_SCALES];
  const float* proj;
  const float* kinv;
} SfmInputs;

/* Gradients w.r.t. the network outputs (base_model.py:59-63 are the producers).  Same shapes as the
 * corresponding inputs.  glogits may be NULL when exp_reg == 0. */
typedef struct SfmGrads {
  float* gdisps[SFM_MAX_SCALES];
  float* gposes;
  float* glogits[SFM_MAX_SCALES];
} SfmGrads;

/* Optional per-(scale, source) dumps of the warp for stage-level parity tests (may be NULL):
 *   P[s]   (B,S,3,h,w) f32  warped image            (transform.py:189)
 *   u0[s]  (B,S,h,w)   i32  floor(u), v0 likewise   (integer work: must be bit-exact)
 *   inb[s] (B,S,h,w)   u8   strict in-bounds flag   (transform.py:128-131)               */
typedef struct SfmDebug {
  float* P[SFM_MAX_SCALES];
  int32_t* u0[SFM_MAX_SCALES];
  int32_t* v0[SFM_MAX_SCALES];
  uint8_t* inb[SFM_MAX_SCALES];
} SfmDebug;

int sfm_version(void);
const char* sfm_last_error(void);

/* Bytes of device scratch a call needs (image pyramid of the scales >= 1, planar like the inputs; projection
 * tables; fp64 reduction cells).  The caller allocates it (any 256-byte aligned device pointer) and passes it to
 * every call with the same descriptor.  Scale 0 is read straight from in->tgt / in->src. */
size_t sfm_workspace_bytes(const SfmDesc* desc);

/* losses_out: device float[5] = total, pixel, smooth, exp, ssim  (chainer.report keys,
 * base_model.py:119-123).  Under snippet sharding they are this shard's partial sums; add across
 * ranks (one allreduce of 5 floats). */

/* Forward only.  Replaces SFMLearner.__call__'s loss loop, base_model.py:64-118
 * (F.resize_images :71-72, projective_inverse_warp transform.py:156-193, compute_smooth_loss :169-185,
 * compute_exp_reg_loss :157-167, compute_ssim :126-142). */
int sfm_loss_forward(const SfmDesc* desc, const SfmInputs* in, float* losses_out, const SfmDebug* debug,
                     void* workspace, void* stream);

/* Backward only (recomputes the forward).  Replaces loss.backward() through the graph recorded by
 * base_model.py:64-118.  gy: device float* holding dL/dloss, or NULL for 1. */
int sfm_loss_backward(const SfmDesc* desc, const SfmInputs* in, const float* gy, const SfmGrads* grads,
                      void* workspace, void* stream);

/* Fused single pass: losses and gradients (for upstream gradient 1) in one sweep.  This is what the
 * Chainer adapter calls from forward(); its backward() only rescales (sfm_scale_grads). */
int sfm_loss_forward_backward(const SfmDesc* desc, const SfmInputs* in, float* losses_out,
                              const SfmGrads* grads, void* workspace, void* stream);

/* grads *= *gy (device scalar).  No-op on the device when *gy == 1. */
int sfm_scale_grads(const SfmDesc* desc, const float* gy, const SfmGrads* grads, void* stream);

/* Profiling hook (bench.py's roofline measurement): while set (non-NULL cudaEvent_t handles), every
 * sfm_loss_* call on this thread records `start` right before and `stop` right after the fused loss
 * kernel, on the call's stream.  Pass NULL, NULL to clear. */
int sfm_set_kernel_events(void* start_event, void* stop_event);

/* Image pyramid only (F.resize_images from full resolution, base_model.py:70-72; scales 1..n_scales-1, scale 0 is
 * the identity and is never copied) into the workspace; sfm_pyramid_export copies one level (scale >= 1) back
 * out for tests: tgt_out (B,3,h,w), src_out (B,S,3,h,w). */
int sfm_pyramid(const SfmDesc* desc, const float* tgt, const float* src, void* workspace, void* stream);
int sfm_pyramid_export(const SfmDesc* desc, const void* workspace, int scale, float* tgt_out, float* src_out,
                       void* stream);

/* Projection tables on the device: proj_out (B,S,n_scales,3,4), kinv_out (B,n_scales,3,3).
 * Replaces proj_tgt_to_src / pose_vec2mat / euler2mat (transform.py:11-91) and F.batch_inv(K) (:105),
 * including the five blocking host<->device copies per (scale, source) of models/utils.py:33-84. */
int sfm_build_tables(const SfmDesc* desc, const float* poses, const float* intrinsics, float* proj_out,
                     float* kinv_out, void* stream);

/* Stage API of the producer-side fusions (the same device code the fused kernels inline; the parity tests
 * use it to show raw-input mode == activation stage + plain mode bit for bit):
 * sfm_disp_activation: disp[k] = DISP_SCALING * sigmoid(x[k]) + MIN_DISP (disp_net.py:104) for n elements,
 * optional dact[k] = d disp / d x.   sfm_pose_reduce: x (B, 6*S, hw) -> poses (B, S, 6) = 0.01 * mean
 * (pose_net.py:52-53). */
int sfm_disp_activation(long long n, const float* x, float* disp, float* dact, void* stream);
int sfm_pose_reduce(int B, int S, int hw, const float* x, float* poses_out, void* stream);

/* Data layer in front of the loss path, as one device gather (datasets/kitti/kitti_raw_dataset.py:12-14
 * load_as_float_norm; datasets/kitti/kitti_raw_transformed.py:23-74 data_augmentation and :76-93
 * get_multi_scale_intrinsics).  The caller draws the random numbers exactly as the reference does and passes
 * them per snippet; NULL `aug` = no augmentation (validation split). */
typedef struct SfmAugment {
  int32_t out_h, out_w;        /* size after random scaling: int(H * y_scaling), int(W * x_scaling)  (:36-37)  */
  int32_t off_y, off_x;        /* random crop offsets in the rescaled image                         (:49-50)  */
  int32_t flip;                /* != 0: horizontal flip                                             (:63-65)  */
  int32_t reserved_;
  double x_scaling, y_scaling; /* np.random.uniform(1, 1.15, 2), used for the intrinsics            (:34-42)  */
} SfmAugment;
/* frames (B, 1+S, H, W, 3) uint8 HWC on the device, frame 0 of a snippet = target (imread order); K_in (B,3,3);
 * aug: device array of B SfmAugment or NULL.  Outputs: tgt_out (B,3,H,W), src_out (B,S,3,H,W) float32 in
 * [-1, 1]; intrinsics_out (B,n_scales,3,3) or NULL. */
int sfm_ingest_u8(int B, int S, int H, int W, int n_scales, const uint8_t* frames, const float* K_in,
                  const SfmAugment* aug, float* tgt_out, float* src_out, float* intrinsics_out, void* stream);

/* Inference-side depth evaluation of one batch (evaluate.py:94-103 + kitti_eval/depth_util.py:6-22), on the
 * device: pred_depth (B,1,h,w) is resized to the ground truth's (Hg,Wg) (F.resize_images), clipped to
 * [min_depth, max_depth], masked, scaled by median(gt)/median(pred) (exact medians) and compared.
 * gt_depth (B,Hg,Wg) float32, mask (B,Hg,Wg) uint8 (non-zero = valid; masked depths must be positive).
 * errors_out: device float[8] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3 (depth_util.py:22) and the scale
 * factor.  scratch: sfm_eval_depth_scratch_bytes(B,Hg,Wg) bytes of device memory.  Asynchronous on `stream`. */
size_t sfm_eval_depth_scratch_bytes(int B, int Hg, int Wg);
int sfm_eval_depth(int B, int h, int w, int Hg, int Wg, const float* pred_depth, const float* gt_depth,
                   const uint8_t* mask, float min_depth, float max_depth, float* errors_out, void* scratch,
                   void* stream);

/* Stage API: projective_inverse_warp(imgs, depthes, poses, K) of transform.py:156-165 on N images of
 * one resolution.  imgs (N,3,h,w) NCHW, depth (N,h*w) [the reference passes it broadcast to 3 rows],
 * poses (N,6), K (N,3,3); proj (N,3,4) / kinv (N,3,3) optional overrides (NULL = built on device).
 * Outputs: out (N,3,h,w); optional u0,v0 (N,h,w) int32 and inb (N,h,w) uint8. */
int sfm_warp_forward(int N, int h, int w, const float* imgs, const float* depth, const float* poses,
                     const float* K, const float* proj, const float* kinv, float* out, int32_t* u0,
                     int32_t* v0, uint8_t* inb, void* stream);
/* Backward of the above for upstream gy (N,3,h,w): gdepth (N,h*w), gposes (N,6), optional gimgs
 * (N,3,h,w; the reference discards it, base_model.py:71-72 `.data`).  scratch: device buffer of
 * sfm_warp_backward_scratch_bytes(N) bytes. */
size_t sfm_warp_backward_scratch_bytes(int N);
int sfm_warp_backward(int N, int h, int w, const float* imgs, const float* depth, const float* poses,
                      const float* K, const float* gy, float* gdepth, float* gposes, float* gimgs,
                      void* scratch, void* stream);

/* SpatialTransformerSamplerInterp (models/spational_transformer_sampler_interp.py:9-159, the repo-local
 * sampler the live path leaves commented out, transform.py:190-191): grid in PIXEL units, indices
 * clamped to the image.  x (B,C,H,W), grid (B,2,oH,oW) -> y (B,C,oH,oW);  backward returns ggrid
 * (B,2,oH,oW) and gx = zeros (:148). */
int sfm_sampler_interp_forward(int B, int C, int H, int W, int oH, int oW, const float* x, const float* grid,
                               float* y, void* stream);
int sfm_sampler_interp_backward(int B, int C, int H, int W, int oH, int oW, const float* x, const float* grid,
                                const float* gy, float* gx, float* ggrid, void* stream);

/* Host-buffer convenience path (what a CPU-resident caller, e.g. a Chainer numpy run or the end-to-end
 * benchmark, uses): owns device buffers, copies inputs H2D, runs sfm_loss_forward_backward, copies
 * losses and gradients D2H and synchronises.  All pointers in `in`, `grads` and `losses_out` are HOST
 * pointers here (pinned memory makes the copies asynchronous). */
typedef struct SfmHostCtx SfmHostCtx;
int sfm_host_ctx_create(const SfmDesc* desc, SfmHostCtx** ctx_out);
int sfm_host_ctx_destroy(SfmHostCtx* ctx);
int sfm_loss_step_host(SfmHostCtx* ctx, const SfmInputs* in, float* losses_out, const SfmGrads* grads);
/* The same step split in two, for callers that keep several steps in flight (a data loader that prepares
 * step k+1 while step k runs): _submit enqueues the H2D copies, the kernels and the D2H copies on the context's
 * own stream and returns; _wait blocks until they are done.  The host buffers of a submitted step must stay
 * valid (and, for the outputs, untouched) until _wait returns.  With two contexts used alternately the copies of
 * one step overlap the kernels of the other (pinned memory required for the overlap).  Host arrays of one kind (disps,
 * logits, gdisps, glogits) that lie back to back in scale order are moved in ONE copy each (every copy has a fixed cost of
 * several microseconds); separate arrays work the same, one copy per scale. */
int sfm_loss_step_host_submit(SfmHostCtx* ctx, const SfmInputs* in, float* losses_out, const SfmGrads* grads);
int sfm_loss_step_host_wait(SfmHostCtx* ctx);

/* The host-buffer step fed by the data layer's raw material: `frames` (B, 1+S, H, W, 3) uint8 HWC, `K_in` (B,3,3)
 * and `aug` (B SfmAugment or NULL) are HOST pointers; the images and intrinsics of `in` are ignored (sfm_ingest_u8
 * produces them on the device), its disps / poses / logits are host pointers as in sfm_loss_step_host_submit.
 * The H2D copy carries 1 byte per image sample instead of 4.  Complete with sfm_loss_step_host_wait. */
int sfm_loss_step_host_u8_submit(SfmHostCtx* ctx, const uint8_t* frames, const float* K_in, const SfmAugment* aug,
                                 const SfmInputs* in, float* losses_out, const SfmGrads* grads);

/* ---- Multi-GPU: the path shards by snippet (SfmDesc.B_global), gradients need no communication, and the five loss
 * partials are completed by ONE all-reduce (sum, float32) per step.  Replaces the reduce of Chainer's
 * MultiprocessParallelUpdater (config_utils.py:123-126, unused by the shipped configs) for this path.
 * One process per GPU.  The communicator wraps an NCCL communicator (bound at run time with dlopen, so the library
 * loads without NCCL; sfm_nccl_set_library names the shared object when the process has not loaded one already):
 *   rank 0:      sfm_comm_unique_id(id)            -> 128 opaque bytes, handed to the other ranks by the host program
 *                                                     (MPI, torch.distributed, a file ...)
 *   every rank:  sfm_comm_create(id, nranks, rank, &comm)     collective, on the rank's current device
 *   every step:  sfm_allreduce_partials(comm, losses, 5, stream)   in place on the device array `losses`; enqueued on
 *                                                     `stream` (after the loss call that wrote it), asynchronous,
 *                                                     capturable in a CUDA graph together with the step
 *   at the end:  sfm_comm_destroy(comm)           after every CUDA graph that captured the all-reduce has been destroyed
 *                                                 (such a graph holds a reference on the NCCL communicator)          */
#define SFM_NCCL_UNIQUE_ID_BYTES 128
typedef struct SfmComm SfmComm;
int sfm_nccl_set_library(const char* path);
int sfm_nccl_version(void);                /* ncclGetVersion code (e.g. 22809), 0 when NCCL cannot be loaded */
int sfm_comm_unique_id(void* id_out);
int sfm_comm_create(const void* unique_id, int nranks, int rank, SfmComm** comm_out);
int sfm_comm_destroy(SfmComm* comm);
int sfm_allreduce_partials(SfmComm* comm, float* losses, int count, void* stream);

/* The same sum WITHOUT a collective call: the epilogue kernel of the step exchanges the five partials over NVLink
 * peer memory itself (every rank owns a small slot array that the other ranks' epilogues write with peer stores; CUDA
 * IPC maps the arrays across the processes) and adds them in rank order, so losses_out holds bitwise the same global
 * sums on every rank when the step's kernels are done -- no extra launch, no library latency on the step.
 *   every rank:  sfm_peer_create(nranks, rank, &peer, handle)   -> 64 opaque bytes (a cudaIpcMemHandle_t)
 *   host side:   all-gather the handles of all ranks in rank order (nranks x 64 bytes)
 *   every rank:  sfm_peer_connect(peer, all_handles); then a host barrier before the first step
 *   every step:  sfm_loss_forward_backward_peer(desc, in, losses_out, grads, workspace, peer, stream)
 *                (graph-capturable; every rank must run the same number of steps: step k of one rank waits, inside the
 *                 epilogue kernel and for at most ~2 s, for step k of the others; on that timeout the losses are NaN)
 *   at the end:  a host barrier, then sfm_peer_destroy(peer)
 * One node, one process per GPU, up to 16 ranks, GPUs with peer access (NVLink / NVSwitch).                        */
#define SFM_IPC_HANDLE_BYTES 64
typedef struct SfmPeer SfmPeer;
int sfm_peer_create(int nranks, int rank, SfmPeer** peer_out, void* ipc_handle_out);
int sfm_peer_connect(SfmPeer* peer, const void* all_handles);
int sfm_peer_destroy(SfmPeer* peer);
int sfm_loss_forward_backward_peer(const SfmDesc* desc, const SfmInputs* in, float* losses_out, const SfmGrads* grads,
                                   void* workspace, SfmPeer* peer, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SFMLOSS_H_ */
