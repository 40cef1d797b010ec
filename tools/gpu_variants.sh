#!/bin/bash
# build the library with different compile-time knobs ON the GPU box and time configs: tools/gpu_variants.sh "cfg1 cfg4" "-DSFM_SI=1 -DSFM_MINB=24" "..."
cfgs=$1; shift
for flags in "$@"; do
  SFM_NVCC_FLAGS="$flags" python sfm_learner_chainer_b200/build.py --force > /dev/null || echo BUILD FAILED
  echo "== $flags"
  timeout 200 python tools/time_kernels.py $cfgs
done
