// ssim_march.cuh -- placeholder
int sfm_launch_ssim(SfmFusedParams& p, bool grad, bool accum, bool debug, long long want_warps, cudaStream_t stream) {
  sfm_set_error("SSIM kernel not built yet");
  return SFM_E_UNSUPPORTED;
}
