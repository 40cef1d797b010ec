#!/bin/bash
# Development aid: builds libsfmloss with extra compile-time knobs into sfm_learner_chainer_b200/variants/lib_<name>.so
# (select it with SFM_LIB_PATH).  usage: tools/build_variant.sh <name> "<nvcc flags>"
set -e
cd "$(dirname "$0")/.."
mkdir -p sfm_learner_chainer_b200/variants
C=sfm_learner_chainer_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $2 -I include -I $C \
  -o sfm_learner_chainer_b200/variants/lib_$1.so $C/api.cu $C/prep.cu $C/smooth.cu $C/fused_loss.cu $C/stage.cu $C/ingest.cu $C/eval.cu $C/comm.cu -ldl
echo built lib_$1.so
