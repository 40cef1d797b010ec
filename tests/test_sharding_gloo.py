"""N>1 host logic on CPU: world_size-2 gloo processes shard the batch by snippet, compute their partial
losses with B_global (the numpy oracle stands in for the CUDA kernels here -- there is no GPU), and the
product's allreduce helper completes the five scalars.  Gradients need no communication."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import sfm_oracle as O
    from sfm_learner_chainer_b200.distributed import shard_range, shard_arrays, allreduce_loss_partials
    from sfm_learner_chainer_b200.synthetic import make_snippets
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B = 5                                                   # uneven split: 3 + 2
    d = make_snippets(B, 2, 32, 104, seed=40)
    mine = shard_arrays(dict(tgt=d['tgt'], src=d['src'], intrinsics=d['intrinsics'], disps=d['disps'],
                             poses=d['poses'], logits=d['logits']), B, rank, world)
    cfg = O.LossConfig(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0, B_global=B)
    L, G, _ = O.sfm_loss(mine['tgt'], mine['src'], mine['intrinsics'], mine['disps'], mine['poses'], mine['logits'], cfg)
    losses = torch.from_numpy(O.losses_vec(L).astype(np.float32))
    work = allreduce_loss_partials(losses, async_op=True)
    work.wait()
    lo, hi = shard_range(B, rank, world)
    # ShardedViewSynthesisLoss without B_global: the global batch is the SUM of the (uneven) local batches, not
    # local * world (3 + 2 = 5, not 6 and 4); an explicit B_global is taken as is; an empty shard is rejected
    from sfm_learner_chainer_b200.distributed import ShardedViewSynthesisLoss
    sh = ShardedViewSynthesisLoss(0.1, 0.2, 0.0, edge_aware_smooth=True)
    sh._resolve_global_batch(torch.from_numpy(mine['src']))
    assert sh.op.B_global == B and sh.op.edge_aware_smooth, (sh.op.B_global, hi - lo)
    sh2 = ShardedViewSynthesisLoss(0.1, 0.2, 0.0, B_global=7)
    sh2._resolve_global_batch(torch.from_numpy(mine['src']))
    assert sh2.op.B_global == 7
    try:
        sh._resolve_global_batch(torch.zeros(0, 2, 3, 32, 104))
        raise AssertionError('empty shard accepted')
    except ValueError:
        pass
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), losses=losses.numpy(), gpose=G['gpose'], lo=lo, hi=hi,
             gdisp0=G['gdisp'][0])
    dist.destroy_process_group()


def test_snippet_sharding_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import sfm_oracle as O
    from sfm_learner_chainer_b200.synthetic import make_snippets
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    d = make_snippets(5, 2, 32, 104, seed=40)
    L, G, _ = O.sfm_loss(d['tgt'], d['src'], d['intrinsics'], d['disps'], d['poses'], d['logits'],
                         O.LossConfig(smooth_reg=0.1, exp_reg=0.2, ssim_rate=0.0))
    r0, r1 = np.load(tmp_path / 'rank0.npz'), np.load(tmp_path / 'rank1.npz')
    assert (int(r0['lo']), int(r0['hi']), int(r1['lo']), int(r1['hi'])) == (0, 3, 3, 5)
    for r in (r0, r1):
        np.testing.assert_allclose(r['losses'], O.losses_vec(L), rtol=2e-6)       # both ranks hold the full-batch losses
        sl = slice(int(r['lo']), int(r['hi']))
        np.testing.assert_allclose(r['gpose'], G['gpose'][sl], rtol=1e-5, atol=1e-9)  # shard gradients are final
        np.testing.assert_allclose(r['gdisp0'], G['gdisp'][0][sl], rtol=1e-5, atol=1e-12)


def test_shard_range_properties():
    from sfm_learner_chainer_b200.distributed import shard_range
    for B in (1, 4, 5, 32, 33, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)
