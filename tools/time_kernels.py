"""Quick device timing of the fused fwd+bwd pass per BASELINE config (development aid, not bench.py)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sfm_learner_chainer_b200 import ViewSynthesisLoss
from sfm_learner_chainer_b200.synthetic import make_snippets, CONFIGS

def algo_bytes(B, S, H, W, exp):
    pix = sum((H >> s) * (W >> s) for s in range(4))
    return 4 * (B * (1 + S) * 3 * H * W + 2 * B * pix + (2 * B * S * pix if exp else 0)) + 48 * B * S + 144 * B

def main():
    names = sys.argv[1:] or ['cfg1', 'cfg2', 'cfg4', 'cfg5']
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for name in names:
        c = dict(CONFIGS[name]); B, S, H, W = c.pop('B'), c.pop('S'), c.pop('H'), c.pop('W')
        nb = min(B, 4)
        d = make_snippets(nb, S, H, W, seed=0)
        rep = lambda a: np.concatenate([a] * (B // nb), 0)
        A = algo_bytes(B, S, H, W, c['exp_reg'] != 0)
        nsets = max(2, int(2 * 128e6 / A) + 1) if A < 256e6 else 2
        sets = []
        for k in range(nsets):
            sets.append(dict(tgt=dev(rep(d['tgt'])), src=dev(rep(d['src'])), K=dev(rep(d['intrinsics'])),
                             disps=[dev(rep(x)) for x in d['disps']], poses=dev(rep(d['poses'])),
                             logits=[dev(rep(x)) for x in d['logits']]))
        op = ViewSynthesisLoss(**c)
        def step(k):
            s = sets[k % nsets]
            return op.forward_backward(s['tgt'], s['src'], s['K'], s['disps'], s['poses'], s['logits'])
        for k in range(5): step(k)
        torch.cuda.synchronize()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n): step(k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        # the fused kernel alone: the library records this event pair immediately around its launch
        import ctypes as C
        from sfm_learner_chainer_b200 import lib as L
        lib = L.load()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a, b in evs:
            a.record(); b.record()
        torch.cuda.synchronize()
        for k, (a, b) in enumerate(evs):
            lib.sfm_set_kernel_events(C.c_void_p(a.cuda_event), C.c_void_p(b.cuda_event))
            step(k)
        lib.sfm_set_kernel_events(None, None)
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        fused_us = 1e3 * ts[len(ts) // 2]
        pix = B * sum((H >> s) * (W >> s) for s in range(4))
        print(json.dumps(dict(cfg=name, ms=round(ms, 4), fused_us=round(fused_us, 1), mpix_s=round(pix / ms / 1e3, 1), algo_MB=round(A / 1e6, 2),
                              gbs=round(A / ms / 1e6, 1), frac_of_6555=round(A / ms / 1e6 / 6555.2, 4), nsets=nsets)))

if __name__ == '__main__':
    main()
