#!/bin/bash
# r2p: raw PCIe bandwidth + ncu full captures of the current kernels
mkdir -p gpurun_out
python - <<'PY'
import torch, time
for mb in (1, 3, 16, 64):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device='cuda')
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    n = 50
    for _ in range(n): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/n
    h2 = torch.empty_like(h).pin_memory()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): h2.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt2=(time.perf_counter()-t)/n
    print('pinned %2d MB: H2D %.1f GB/s  D2H %.1f GB/s' % (mb, (mb<<20)/dt/1e9, (mb<<20)/dt2/1e9))
PY
for cfg in cfg2 cfg4 cfg5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 3 -c 1 -f -o gpurun_out/r2p_fused_$cfg python tools/time_kernels.py $cfg > gpurun_out/r2p_ncu_$cfg.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:prep -s 3 -c 1 -f -o gpurun_out/r2p_prologue_cfg5 python tools/time_kernels.py cfg5 > gpurun_out/r2p_ncu_prep.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:prep -s 3 -c 1 -f -o gpurun_out/r2p_prologue_cfg4 python tools/time_kernels.py cfg4 > gpurun_out/r2p_ncu_prep4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail
