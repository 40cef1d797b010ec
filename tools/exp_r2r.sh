#!/bin/bash
for v in m16 m10; do
  echo "== $v"
  SFM_LIB_PATH=$PWD/sfm_learner_chainer_b200/variants/lib_$v.so timeout 300 python tools/exp_variants.py cfg2 cfg5 -- "" 2>&1
done
