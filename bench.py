#!/usr/bin/env python
"""bench.py -- fwd+bwd warp+photometric loss throughput of the view-synthesis loss path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (image pyramid + fused loss forward+backward, single sweep)
over one batch of synthetic KITTI-shaped snippets.  Default workload = BASELINE.json configs[1]
(sfm_learner_v1_ssim.yml loss path: SSIM + L1 + smoothness, B=4, S=2, 128x416, 4 scales) per GPU;
under N GPUs the batch is sharded by snippet (weak scaling: 4 snippets per rank, B_global = 4N), the
only exchange being the 5-float loss-partial allreduce (asynchronous, NCCL).

Prints ONE JSON line (rank 0).  Keys beyond the driver's contract:
  roofline       dominant kernel (fused loss) vs the measured HBM copy peak, timed live with CUDA events
  roofline_step  whole step vs the strict byte model (SURVEY 8(d) "A-strict")
  cpu_baseline   the numpy oracle (port of the reference's numpy/Chainer CPU path) on this box's cores
  e2e            same metric through the C-ABI host-buffer entry point (H2D + kernels + D2H per step)
  other_configs  device-timed numbers for the remaining single-GPU BASELINE shapes (cfg1, cfg4, cfg5)
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

from sfm_learner_chainer_b200.synthetic import make_snippets, CONFIGS

METRIC = 'fwd+bwd warp+photometric loss Mpix/s at 128x416'
UNIT = 'Mpix/s'
FALLBACK_HBM_GBS = 6650.0


def pyramid_pixels(H, W, n_scales=4):
    return sum((H >> s) * (W >> s) for s in range(n_scales))


def bytes_strict(B, S, H, W, exp):
    """SURVEY 8(d) A-strict: full-res images once + disp r/w + logits r/w + poses + K."""
    pix = pyramid_pixels(H, W)
    return 4 * (B * (1 + S) * 3 * H * W + 2 * B * pix + (2 * B * S * pix if exp else 0)) + 48 * B * S + 144 * B


def bytes_fused_kernel(B, S, H, W, exp):
    """SURVEY 8(d) secondary model (pyramid as the kernel's input): 4*B*sum_hw*(3 + 3S + 2 + exp*2S)."""
    return 4 * B * pyramid_pixels(H, W) * (3 + 3 * S + 2 + (2 * S if exp else 0))


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ---------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the numpy oracle on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_oracle_run(cfg_name, n_iters, threads):
    """Times fwd+bwd of the numpy restatement of the reference path, sharded by snippet over a thread
    pool (numpy releases the GIL in its array loops).  Returns seconds per step (all B snippets)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import sfm_oracle as O                                  # the timed CPU baseline
    c = dict(CONFIGS[cfg_name])
    B, S, H, W = c.pop('B'), c.pop('S'), c.pop('H'), c.pop('W')
    d = make_snippets(B, S, H, W, seed=0)
    cfg = O.LossConfig(B_global=B, **c)

    def one(b):
        sl = slice(b, b + 1)
        return O.sfm_loss(d['tgt'][sl], d['src'][sl], d['intrinsics'][sl], [x[sl] for x in d['disps']],
                          d['poses'][sl], [x[sl] for x in d['logits']], cfg)[0]

    # every (step, snippet) pair is an independent unit of host work: all of them go through one thread pool so
    # that the arm uses every host core it can even when B is smaller than the core count
    threads = max(1, min(threads, B * max(1, n_iters)))
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, range(min(B, threads))))                        # warm-up (page-in, caches)
        t0 = time.perf_counter()
        list(ex.map(one, [b for _ in range(n_iters) for b in range(B)]))
        sec = (time.perf_counter() - t0) / n_iters
    return float(sec), threads, (B, S, H, W)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    c = CONFIGS[args.config]
    sec, threads, (B, S, H, W) = cpu_oracle_run(args.config, max(1, args.steps), ncpu)
    pix = B * pyramid_pixels(H, W)
    val = pix / sec / 1e6
    sample = '%d full %s steps (B=%d, S=%d, %dx%d, 4 scales), (step, snippet) units over a pool of %d threads' % (
        max(1, args.steps), args.config, B, S, H, W, threads)
    line = dict(impl='reference', metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=max(1, args.steps),
                warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload='%s: %s' % (args.config, describe(args.config)), B=B, S=S, H=H, W=W,
                            note='numpy restatement (oracle/) of the reference numpy/Chainer CPU path; '
                                 'chainer==4.0.0b1 is not installable in this image'),
                cpu_baseline=dict(value=val, unit=UNIT, cores=threads, kind='port', sample=sample),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


def describe(name):
    c = CONFIGS[name]
    return 'B=%d S=%d %dx%d 4 scales, smooth_reg=%g exp_reg=%g ssim_rate=%g' % (
        c['B'], c['S'], c['H'], c['W'], c['smooth_reg'], c['exp_reg'], c['ssim_rate'])


# ---------------------------------------------------------------------------------------------
# clocks (pynvml sampler thread)
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index):
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                                          # noqa: BLE001
            self.err = repr(e)
            self.max_mhz = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _sample(self):
        nv = self.nv
        return (time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self._sample())
            except Exception:                                           # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self, t0, t1):
        if not self.ok:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvml unavailable: %s' % getattr(self, 'err', '')])
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or [s for s in self.samples if s[0] >= t0][:1] \
            or self.samples[-1:]
        mhz = float(np.median([s[1] for s in inside])) if inside else None
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [n for b, n in self.REASONS.items() if mask & b and n != 'gpu_idle']
        return dict(sm_mhz=mhz, sm_max_mhz=float(self.max_mhz), reasons=reasons, samples=len(inside))


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class Workload(object):
    """Pre-packed C-ABI calls over `nsets` rotating buffer sets (inputs AND outputs), so that with a
    working set below the 126 MB L2 consecutive steps still touch HBM."""

    def __init__(self, cfg_name, device, B_global=None, min_rotation_bytes=None, B_local=None, seed=0):
        import torch
        from sfm_learner_chainer_b200 import lib as L
        self.torch, self.L = torch, L
        self.lib = L.load()
        c = dict(CONFIGS[cfg_name])
        B, S, H, W = c.pop('B'), c.pop('S'), c.pop('H'), c.pop('W')
        if B_local:
            B = B_local
        self.B, self.S, self.H, self.W, self.flags = B, S, H, W, c
        self.exp = c['exp_reg'] != 0
        self.A_strict = bytes_strict(B, S, H, W, self.exp)
        self.A_kernel = bytes_fused_kernel(B, S, H, W, self.exp)
        self.pix = B * pyramid_pixels(H, W)
        l2 = torch.cuda.get_device_properties(device).L2_cache_size
        rot = min_rotation_bytes if min_rotation_bytes is not None else 2 * l2
        self.nsets = max(2, int(math.ceil(rot / float(self.A_strict))) + 1)
        self.l2_bytes = l2
        nb = min(B, 4)
        base = make_snippets(nb, S, H, W, seed=seed)
        rep = lambda a: np.ascontiguousarray(np.concatenate([a] * (B // nb) + [a[:B % nb]], 0)) if B != nb else a
        self.host = dict(tgt=rep(base['tgt']), src=rep(base['src']), intrinsics=rep(base['intrinsics']),
                         disps=[rep(x) for x in base['disps']], poses=rep(base['poses']),
                         logits=[rep(x) for x in base['logits']])
        self.desc = L.SfmDesc(B, S, H, W, 4, int(B_global or 0), c['smooth_reg'], c['exp_reg'], c['ssim_rate'], 0)
        nws = self.lib.sfm_workspace_bytes(C.byref(self.desc))
        dev = lambda a: torch.from_numpy(a).to(device)
        self.sets = []
        # the loss partials of all buffer sets live in one tensor (row k = set k) so that the multi-GPU run can
        # all-reduce the partials of `nsets` consecutive steps with one NCCL call
        self.all_losses = torch.zeros(self.nsets, 8, device=device)
        for k in range(self.nsets):
            t = dict(tgt=dev(self.host['tgt']), src=dev(self.host['src']), K=dev(self.host['intrinsics']),
                     disps=[dev(x) for x in self.host['disps']], poses=dev(self.host['poses']),
                     logits=[dev(x) for x in self.host['logits']] if self.exp else None,
                     gdisps=[torch.empty_like(dev(x)) for x in self.host['disps']],
                     gposes=torch.empty((B, S, 6), device=device),
                     glogits=[torch.empty((B, S, H >> s, W >> s), device=device) for s in range(4)] if self.exp else None,
                     losses=self.all_losses[k],
                     ws=torch.empty(nws + 256, dtype=torch.uint8, device=device))
            inp, g = L.SfmInputs(), L.SfmGrads()
            inp.tgt, inp.src, inp.intrinsics, inp.poses = (t['tgt'].data_ptr(), t['src'].data_ptr(), t['K'].data_ptr(),
                                                           t['poses'].data_ptr())
            g.gposes = t['gposes'].data_ptr()
            for s in range(4):
                inp.disps[s] = t['disps'][s].data_ptr()
                g.gdisps[s] = t['gdisps'][s].data_ptr()
                if self.exp:
                    inp.logits[s] = t['logits'][s].data_ptr()
                    g.glogits[s] = t['glogits'][s].data_ptr()
            t['inp'], t['g'] = inp, g
            t['wsp'] = C.c_void_p((t['ws'].data_ptr() + 255) // 256 * 256)
            self.sets.append(t)
        self.graphs = None

    def launch(self, k, stream):
        t = self.sets[k % self.nsets]
        rc = self.lib.sfm_loss_forward_backward(C.byref(self.desc), C.byref(t['inp']), C.c_void_p(t['losses'].data_ptr()),
                                                C.byref(t['g']), t['wsp'], C.c_void_p(stream))
        if rc:
            self.L.check(rc)

    def capture(self):
        """One CUDA graph per buffer set (prep + fused kernel nodes): replays cost one launch each."""
        torch = self.torch
        self.graphs = []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(self.nsets):
                self.launch(k, side.cuda_stream)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for k in range(self.nsets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self.launch(k, torch.cuda.current_stream().cuda_stream)
            self.graphs.append(g)
        torch.cuda.synchronize()

    def time_pipelined(self, steps, warmup):
        """Throughput with the image pyramid of batch k+1 built on a side stream while the loss kernels of batch k
        run.  The pyramid (sfm_pyramid) depends on the input images alone -- in a training step it can be issued
        as soon as the batch is on the device, long before the CNN outputs exist -- so the dependent part of the
        path is sfm_loss_forward_backward with SFM_FLAG_REUSE_PYRAMID (tables, smoothness, fused loss, epilogue).
        Every step still does all of its work inside the timed region.  Returns ms per step."""
        torch, L = self.torch, self.L
        desc_r = L.SfmDesc(self.desc.B, self.desc.S, self.desc.H, self.desc.W, 4, self.desc.B_global, self.desc.smooth_reg,
                           self.desc.exp_reg, self.desc.ssim_rate, L.SFM_FLAG_REUSE_PYRAMID)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

        def pyr(k, stream):
            t = self.sets[k]
            L.check(self.lib.sfm_pyramid(C.byref(self.desc), C.c_void_p(t['tgt'].data_ptr()), C.c_void_p(t['src'].data_ptr()),
                                         t['wsp'], C.c_void_p(stream)))

        def loss(k, stream):
            t = self.sets[k]
            L.check(self.lib.sfm_loss_forward_backward(C.byref(desc_r), C.byref(t['inp']), C.c_void_p(t['losses'].data_ptr()),
                                                       C.byref(t['g']), t['wsp'], C.c_void_p(stream)))
        torch.cuda.synchronize()
        with torch.cuda.stream(sa):
            for k in range(self.nsets):
                pyr(k, sa.cuda_stream)
                loss(k, sa.cuda_stream)
        torch.cuda.synchronize()
        gp, gl = [], []
        for k in range(self.nsets):
            a, b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(a, stream=sb):
                pyr(k, torch.cuda.current_stream().cuda_stream)
            with torch.cuda.graph(b, stream=sa):
                loss(k, torch.cuda.current_stream().cuda_stream)
            gp.append(a)
            gl.append(b)
        torch.cuda.synchronize()
        ev_p = [torch.cuda.Event() for _ in range(self.nsets)]
        ev_l = [torch.cuda.Event() for _ in range(self.nsets)]
        used = [False] * self.nsets
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sb):
            gp[0].replay()
            ev_p[0].record(sb)
        for k in range(warmup + steps):
            j, jn = k % self.nsets, (k + 1) % self.nsets
            if k == warmup:
                sa.synchronize()
                sb.synchronize()
                e0.record(sa)
            with torch.cuda.stream(sb):                       # pyramid of the next batch
                if used[jn]:
                    sb.wait_event(ev_l[jn])                   # its workspace is free again
                gp[jn].replay()
                ev_p[jn].record(sb)
            with torch.cuda.stream(sa):                       # loss of this batch
                sa.wait_event(ev_p[j])
                gl[j].replay()
                ev_l[j].record(sa)
                used[j] = True
        e1.record(sa)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def step(self, k):
        if self.graphs is not None:
            self.graphs[k % self.nsets].replay()
        else:
            self.launch(k, self.torch.cuda.current_stream().cuda_stream)

    def time_steps(self, steps, warmup, per_step=None):
        torch = self.torch
        for k in range(warmup):
            self.step(k)
            if per_step:
                per_step(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for k in range(steps):
            self.step(warmup + k)
            if per_step:
                per_step(warmup + k)
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        return e0.elapsed_time(e1) / steps, t0, t1

    def time_fused_kernel(self, n=200):
        """Average device duration of the fused loss kernel alone (events recorded by the library right
        around its launch, on the launching stream)."""
        torch = self.torch
        st = torch.cuda.current_stream()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for k in range(5):
            self.launch(k, st.cuda_stream)
        torch.cuda.synchronize()
        # torch creates its cudaEvent_t lazily on first record(); record once so the handles exist
        for a, b in evs:
            a.record(); b.record()
        torch.cuda.synchronize()
        for k, (a, b) in enumerate(evs):
            self.lib.sfm_set_kernel_events(C.c_void_p(a.cuda_event), C.c_void_p(b.cuda_event))
            self.launch(k, st.cuda_stream)
        self.lib.sfm_set_kernel_events(None, None)
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        return float(np.mean(ts)), float(ts[len(ts) // 2])


def run_e2e(cfg_name, steps, device, n_ctx=3, u8=False):
    """Same metric through the host-buffer C-ABI entry point: pinned host inputs -> H2D -> pyramid + fused
    fwd+bwd -> D2H of the five losses and every gradient, EVERY step.  `n_ctx` host contexts are used alternately
    (sfm_loss_step_host_submit / _wait), the way a data loader keeps the next step's copies in flight while the
    current one computes; n_ctx=1 is the fully synchronous call.  Returns (Mpix/s, h2d bytes, d2h bytes, loss)."""
    import torch
    from sfm_learner_chainer_b200 import lib as L
    lib = L.load()
    c = dict(CONFIGS[cfg_name])
    B, S, H, W = c.pop('B'), c.pop('S'), c.pop('H'), c.pop('W')
    exp = c['exp_reg'] != 0
    d = make_snippets(B, S, H, W, seed=1)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hin = dict(tgt=pin(d['tgt']), src=pin(d['src']), K=pin(d['intrinsics']), poses=pin(d['poses']),
               disps=[pin(x) for x in d['disps']], logits=[pin(x) for x in d['logits']])
    desc = L.SfmDesc(B, S, H, W, 4, 0, c['smooth_reg'], c['exp_reg'], c['ssim_rate'], 0)
    inp = L.SfmInputs()
    inp.tgt, inp.src, inp.intrinsics, inp.poses = hin['tgt'].data_ptr(), hin['src'].data_ptr(), hin['K'].data_ptr(), hin['poses'].data_ptr()
    h2d = hin['tgt'].numel() * 4 + hin['src'].numel() * 4 + hin['K'].numel() * 4 + hin['poses'].numel() * 4
    d2h = 5 * 4 + hin['poses'].numel() * 4
    for s in range(4):
        inp.disps[s] = hin['disps'][s].data_ptr()
        h2d += hin['disps'][s].numel() * 4
        d2h += hin['disps'][s].numel() * 4
        if exp:
            inp.logits[s] = hin['logits'][s].data_ptr()
            h2d += hin['logits'][s].numel() * 4
            d2h += hin['logits'][s].numel() * 4
    if u8:
        # data-layer entry: decoded uint8 frames + augmentation draws instead of float images (sfm_ingest_u8 on the device)
        from sfm_learner_chainer_b200.functions import draw_augmentation
        rs = np.random.RandomState(2)
        imgs = np.concatenate([d['tgt'][:, None], d['src']], 1)
        frames = pin(np.clip(np.round((imgs.transpose(0, 1, 3, 4, 2) + 1) * 127.5), 0, 255).astype(np.uint8))
        K0 = pin(d['intrinsics'][:, 0].copy())
        aug = (L.SfmAugment * B)()
        for b in range(B):
            a = draw_augmentation(H, W, rs)
            aug[b] = L.SfmAugment(a['out_h'], a['out_w'], a['off_y'], a['off_x'], 1 if a['flip'] else 0, 0, a['x_scaling'], a['y_scaling'])
        h2d += frames.numel() + K0.numel() * 4 + C.sizeof(aug) - (hin['tgt'].numel() + hin['src'].numel() + hin['K'].numel()) * 4
    ctxs, outs, grads = [], [], []
    try:
        for k in range(n_ctx):
            ctx = C.c_void_p()
            L.check(lib.sfm_host_ctx_create(C.byref(desc), C.byref(ctx)))
            ctxs.append(ctx)
            ho = dict(gdisps=[pin(np.empty_like(x)) for x in d['disps']], glogits=[pin(np.empty_like(x)) for x in d['logits']],
                      gposes=pin(np.empty_like(d['poses'])), losses=pin(np.zeros(8, np.float32)))
            g = L.SfmGrads()
            g.gposes = ho['gposes'].data_ptr()
            for s in range(4):
                g.gdisps[s] = ho['gdisps'][s].data_ptr()
                if exp:
                    g.glogits[s] = ho['glogits'][s].data_ptr()
            outs.append(ho)
            grads.append(g)

        def run(n):
            for k in range(n):
                j = k % n_ctx
                if k >= n_ctx:
                    L.check(lib.sfm_loss_step_host_wait(ctxs[j]))      # the step submitted n_ctx steps ago: its results are on the host
                if u8:
                    L.check(lib.sfm_loss_step_host_u8_submit(ctxs[j], C.c_void_p(frames.data_ptr()), C.c_void_p(K0.data_ptr()), aug,
                                                             C.byref(inp), C.c_void_p(outs[j]['losses'].data_ptr()), C.byref(grads[j])))
                else:
                    L.check(lib.sfm_loss_step_host_submit(ctxs[j], C.byref(inp), C.c_void_p(outs[j]['losses'].data_ptr()), C.byref(grads[j])))
            for j in range(n_ctx):
                L.check(lib.sfm_loss_step_host_wait(ctxs[j]))
        run(3 * n_ctx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(steps)
        sec = (time.perf_counter() - t0) / steps
        loss = float(outs[(steps - 1) % n_ctx]['losses'][0])
    finally:
        for ctx in ctxs:
            lib.sfm_host_ctx_destroy(ctx)
    return B * pyramid_pixels(H, W) / sec / 1e6, h2d, d2h, loss


def run_b200(args):
    import torch
    import torch.distributed as dist
    from sfm_learner_chainer_b200.distributed import allreduce_loss_partials
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the view-synthesis loss path has no CPU fallback)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    n_gpus = world
    peak, peak_src = measured_peak()

    c = CONFIGS[args.config]
    wl = Workload(args.config, device, B_global=c['B'] * world)
    if not args.no_graph:
        wl.capture()

    # Loss partials (the path's only cross-rank quantity, reporting only -- gradients are final per rank): every
    # step's five partials are summed over the ranks.  --allreduce-every 1 issues one NCCL call per step; the
    # default batches the partials of `nsets` consecutive steps (one row per step, snapshot first so that later
    # steps can overwrite their rows) into one asynchronous call, the way a trainer that reports every few
    # iterations would.  At ~48 us per step the per-step call is host-bound (measured at N=2: 52.0 vs 47.1 us).
    pending = []
    every = args.allreduce_every if args.allreduce_every > 0 else max(1, min(wl.nsets, args.steps))   # >= 1 call inside the timed region
    staging = [torch.zeros_like(wl.all_losses) for _ in range(2)]
    flip = [0]

    def per_step(k):
        if world == 1 or args.no_allreduce:
            return
        if every == 1:
            pending.append(allreduce_loss_partials(wl.sets[k % wl.nsets]['losses'][:5], async_op=True))
        elif (k + 1) % every == 0:
            buf = staging[flip[0]]
            flip[0] ^= 1
            buf.copy_(wl.all_losses)             # ordered after the steps that wrote the rows (same stream)
            pending.append(allreduce_loss_partials(buf, async_op=True))
        if len(pending) > 2 and every > 1 or len(pending) > 64:
            pending.pop(0).wait()

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, t0, t1 = wl.time_steps(args.steps, args.warmup, per_step if world > 1 else None)
    for w in pending:
        w.wait()
    torch.cuda.synchronize()
    allreduce_ok = None
    if world > 1 and not args.no_allreduce and every > 1 and pending:
        # every rank holds the same synthetic snippets, so the reduced partials must be world x the local ones
        red = staging[flip[0] ^ 1][:, :5]
        allreduce_ok = bool(torch.allclose(red, wl.all_losses[:, :5] * world, rtol=1e-5, atol=1e-8))
    if world > 1:
        dist.barrier()
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # keep the GPU busy a little longer if the timed region was too short for a clock sample
    if t1 - t0 < 0.05:
        tt = time.perf_counter()
        k = 0
        while time.perf_counter() - tt < 0.1:
            wl.step(k)
            k += 1
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    sampler.stop()
    clocks = sampler.summary(t0, t1)

    total_pix = wl.pix * world
    value = total_pix / (ms * 1e-3) / 1e6
    n_launch = 3 + (1 if c['smooth_reg'] else 0)       # prep, [smooth], fused, epilogue
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=dict(workload='%s: %s' % (args.config, describe(args.config)), per_gpu_batch=wl.B,
                            global_batch=wl.B * world, sources=wl.S, H=wl.H, W=wl.W, n_scales=4,
                            parallelism=('snippet-sharded x%d, no data-path collective; loss partials of every step all-reduced '
                                         'asynchronously over NCCL, %s' % (world, 'one call per step' if every == 1 else
                                                                           '%d steps per call' % every)) if world > 1 else 'single GPU',
                            l2_policy='inputs+outputs rotated over %d buffer sets (%.0f MB > 2 x L2 %.0f MB)' % (
                                wl.nsets, wl.nsets * wl.A_strict / 1e6, wl.l2_bytes / 1e6),
                            launch=('CUDA graph replay of the step\'s %d kernel nodes (pyramid/tables, %sfused loss, epilogue; '
                                    'programmatic dependent launches between them)' % (n_launch, 'smoothness, ' if c['smooth_reg'] else ''))
                            if not args.no_graph else 'direct C-ABI calls',
                            units='target-pyramid pixels = B * sum_s h_s*w_s (%d per step per GPU)' % wl.pix),
                clocks=clocks, gpu_launches=n_launch * args.steps)
    if allreduce_ok is not None:
        line['loss_allreduce_check'] = allreduce_ok

    if rank == 0:
        # ---- roofline of the dominant kernel (fused loss), events around the kernel itself
        k_mean, k_med = wl.time_fused_kernel(200)
        ach = wl.A_kernel / (k_mean * 1e-3) / 1e9
        traffic, issue = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_latest.json')))[args.config]
            traffic = prof.get('dram_bytes')             # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
            issue = prof
        except Exception:
            pass
        line['roofline'] = dict(bound='hbm', achieved=ach, peak=peak, unit='GB/s', frac=ach / peak, traffic=traffic,
                                kernel='sfm_fused_%s_kernel' % ('ssim' if (c['ssim_rate'] and not c['exp_reg']) else 'l1'),
                                kernel_us=k_mean * 1e3, kernel_us_median=k_med * 1e3,
                                algorithmic_bytes=wl.A_kernel, peak_source=peak_src,
                                byte_model='4*B*sum_hw*(3 + 3S + 2 + exp*2S): NHWC4 pyramid in, disp in, gdisp out (+logits/glogits)',
                                note='at B=4 the roofline time is ~2 us (below launch latency): latency-bound; '
                                     'see other_configs for the bandwidth-relevant shapes.  The fused kernels are '
                                     'FP32-issue bound, not HBM bound (DESIGN.md section 5): ncu of the same kernel in '
                                     'profiles/ncu_latest.json',
                                ncu=issue)
        ach_s = wl.A_strict / (ms * 1e-3) / 1e9
        line['roofline_step'] = dict(bound='hbm', achieved=ach_s, peak=peak, unit='GB/s', frac=ach_s / peak,
                                     algorithmic_bytes=wl.A_strict,
                                     byte_model='A-strict: full-res images once + disp r/w + logits r/w + poses + K')
        line['step_share_of_fused_kernel'] = k_mean / ms if world == 1 else None
    if world == 1:
        # ---- the same steps with the next batch's pyramid overlapped (extra figure; `value` stays the sequential step)
        if not args.no_other and not args.no_graph:
            try:
                pms = wl.time_pipelined(min(args.steps, 1000), args.warmup)
                line['pipelined'] = dict(value=wl.pix / (pms * 1e-3) / 1e6, unit=UNIT, ms_per_step=pms,
                                         note='sfm_pyramid of batch k+1 on a side stream while sfm_loss_forward_backward('
                                              'SFM_FLAG_REUSE_PYRAMID) of batch k runs; the pyramid depends on the input images only')
            except Exception as exc:                          # noqa: BLE001 -- an extra figure must not take the line down
                line['pipelined'] = dict(error=str(exc)[:200])
        # ---- e2e through the host-buffer C-ABI entry point
        e2e_steps = max(5, min(200, args.steps))
        ev, h2d, d2h, _ = run_e2e(args.config, e2e_steps, device, n_ctx=3)
        ev1, _, _, _ = run_e2e(args.config, e2e_steps, device, n_ctx=1)
        ev8, h2d8, _, _ = run_e2e(args.config, e2e_steps, device, n_ctx=3, u8=True)
        line['e2e'] = dict(value=ev, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps,
                           api='sfm_loss_step_host_submit/_wait, three host contexts used in rotation (pinned host buffers; every step: '
                               'H2D of all inputs, pyramid + fused fwd+bwd, D2H of losses and all gradients)',
                           synchronous_value=ev1, synchronous_api='sfm_loss_step_host (one step at a time)',
                           u8_frames_value=ev8, u8_frames_h2d_bytes_per_step=h2d8,
                           u8_frames_api='sfm_loss_step_host_u8_submit/_wait: decoded uint8 frames + augmentation draws in, '
                                         'normalisation / random scale-crop-flip / multi-scale intrinsics on the device (data-layer fusion)')
        # ---- other single-GPU BASELINE shapes, device timed
        if not args.no_other:
            others = {}
            del wl
            torch.cuda.empty_cache()
            for name in ('cfg1', 'cfg4', 'cfg5'):
                if name == args.config:
                    continue
                w2 = Workload(name, device)
                w2.capture()
                oms, _, _ = w2.time_steps(50, 5)
                km, _ = w2.time_fused_kernel(50)
                others[name] = dict(workload=describe(name), ms_per_step=oms, value=w2.pix / (oms * 1e-3) / 1e6, unit=UNIT,
                                    step_frac_of_hbm_peak=w2.A_strict / (oms * 1e-3) / 1e9 / peak,
                                    fused_kernel_us=km * 1e3,
                                    fused_kernel_frac_of_hbm_peak=w2.A_kernel / (km * 1e-3) / 1e9 / peak)
                del w2
                torch.cuda.empty_cache()
            line['other_configs'] = others
        # ---- CPU baseline: bounded sample of the same workload on the host cores
        if not args.no_cpu:
            n_iters = 12 if args.config in ('cfg1', 'cfg2') else 1
            sec, threads, (B, S, H, W) = cpu_oracle_run(args.config, n_iters, 1)
            line['cpu_baseline'] = dict(value=B * pyramid_pixels(H, W) / sec / 1e6, unit=UNIT, cores=threads, kind='port',
                                        sample='%d full %s steps of the numpy oracle (scalar port, 1 thread), median' % (n_iters, args.config),
                                        ms_per_step=sec * 1e3)
    else:
        # e2e at N GPUs: every rank runs the host-buffer path on its shard; aggregate = sum / max time
        e2e_steps = max(5, min(100, args.steps))
        dist.barrier()
        ev, h2d, d2h, _ = run_e2e(args.config, e2e_steps, device)
        t = torch.tensor([wl.pix / ev], device=device)      # us per step on this rank (pix / Mpix/s)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line['e2e'] = dict(value=total_pix / float(t.item()), unit=UNIT, h2d_bytes_per_step=h2d * world,
                           d2h_bytes_per_step=d2h * world, steps=e2e_steps,
                           api='sfm_loss_step_host on every rank (its snippet shard)')
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=sorted(CONFIGS))
    ap.add_argument('--no-graph', action='store_true', help='direct C-ABI calls instead of CUDA graph replay')
    ap.add_argument('--no-other', action='store_true', help='skip the other_configs sweep')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-allreduce', action='store_true', help='(diagnostic) skip the loss-partial allreduce at N > 1')
    ap.add_argument('--allreduce-every', type=int, default=0,
                    help='N > 1: all-reduce the loss partials every K steps (K rows in one call); 1 = one call per step; '
                         '0 (default) = one call per rotation of the buffer sets')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        if args.steps > 20:
            args.steps = 20            # bounded sample: each step is ~0.2-0.8 s of host work
        if args.steps < 8:
            args.steps = 8             # enough (step, snippet) units to occupy the host cores
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
