"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
SURVEY T6 -- the G-GPU snippet-sharded result equals the 1-GPU result.  One process per GPU over NCCL, each rank runs
ShardedViewSynthesisLoss on its block of snippets with B_global; its gradients must be final (no communication) and
the all-reduced loss partials must equal the full-batch losses."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = dict(smooth_reg=0.1, exp_reg=0.0, ssim_rate=0.15)


def _worker(rank, world, port, out_dir, B):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sfm_learner_chainer_b200.distributed import ShardedViewSynthesisLoss, shard_arrays, shard_range
    from sfm_learner_chainer_b200.synthetic import make_snippets
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    d = make_snippets(B, 2, 128, 416, seed=90)
    mine = shard_arrays(dict(tgt=d['tgt'], src=d['src'], intrinsics=d['intrinsics'], disps=d['disps'], poses=d['poses'],
                             logits=d['logits']), B, rank, world)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    op = ShardedViewSynthesisLoss(B_global=B, **FLAGS)
    losses, grads, work = op.forward_backward(dev(mine['tgt']), dev(mine['src']), dev(mine['intrinsics']),
                                              [dev(x) for x in mine['disps']], dev(mine['poses']), None, async_op=True)
    work.wait()
    torch.cuda.synchronize()
    # the same through the C ABI's own communicator (sfm_comm_* / sfm_allreduce_partials), captured in a CUDA graph
    # together with the step: uneven shards, B_global left to the operator (sum of the local batches)
    from sfm_learner_chainer_b200.distributed import LossPartialsComm
    comm = LossPartialsComm(rank, world)
    op2 = ShardedViewSynthesisLoss(comm=comm, **FLAGS)
    args = (dev(mine['tgt']), dev(mine['src']), dev(mine['intrinsics']), [dev(x) for x in mine['disps']], dev(mine['poses']), None)
    l2, g2, _ = op2.forward_backward(*args)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        l3, g3, _ = op2.forward_backward(*args)            # warm-up on the capture stream
        side.synchronize()
        # thread-local capture mode: torch's NCCL watchdog thread may query events while this thread captures
        with torch.cuda.graph(graph, stream=side, capture_error_mode='thread_local'):
            l3, g3, _ = op2.forward_backward(*args)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    l3_host = l3.cpu().numpy()
    # ... and with NO collective call: the epilogue kernel sums the partials over NVLink peer memory (sfm_peer_*)
    from sfm_learner_chainer_b200.distributed import PeerLossSum
    peer = PeerLossSum(rank, world)
    op3 = ShardedViewSynthesisLoss(peer=peer, **FLAGS)
    for _ in range(3):
        l4, g4, _ = op3.forward_backward(*args)
    torch.cuda.synchronize()
    graph2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        op3.forward_backward(*args)
        side.synchronize()
        with torch.cuda.graph(graph2, stream=side, capture_error_mode='thread_local'):
            l5, _, _ = op3.forward_backward(*args)
    for _ in range(5):
        graph2.replay()
    torch.cuda.synchronize()
    l4_host, l5_host, gd4 = l4.cpu().numpy(), l5.cpu().numpy(), g4['gdisps'][0].cpu().numpy()
    del graph2
    peer.close()
    # a graph that captured the all-reduce holds a reference on the NCCL communicator: destroy it before the communicator
    # (ncclCommDestroy waits for such references)
    del graph, g3
    import gc
    gc.collect()
    lo, hi = shard_range(B, rank, world)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), losses=losses.cpu().numpy(), gposes=grads['gposes'].cpu().numpy(),
             gdisp0=grads['gdisps'][0].cpu().numpy(), lo=lo, hi=hi, losses_abi=l2.cpu().numpy(), losses_graph=l3_host,
             gdisp0_abi=g2['gdisps'][0].cpu().numpy(), B_global=op2.op.B_global, losses_peer=l4_host, losses_peer_graph=l5_host,
             gdisp0_peer=gd4)
    comm.close()
    dist.destroy_process_group()


def test_sharded_result_equals_single_gpu_result(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    import torch.multiprocessing as mp
    from sfm_learner_chainer_b200 import ViewSynthesisLoss
    from sfm_learner_chainer_b200.synthetic import make_snippets
    from tests.gpu_util import dev_inputs, host, assert_grad_close
    world, B = 2, 5                                   # uneven shards: 3 + 2
    port = 29600 + (os.getpid() % 2000)
    ctx = mp.spawn(_worker, args=(world, port, str(tmp_path), B), nprocs=world, join=False)
    import time
    deadline = time.time() + 240                       # a stuck collective must not eat the GPU budget
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for pr in ctx.processes:
                pr.kill()
            pytest.fail('multi-GPU workers did not finish within 240 s')
    d = make_snippets(B, 2, 128, 416, seed=90)
    g = dev_inputs(d)
    lf, gf = ViewSynthesisLoss(**FLAGS).forward_backward(g['tgt'], g['src'], g['intrinsics'], g['disps'], g['poses'], None)
    for r in range(world):
        z = np.load(tmp_path / ('rank%d.npz' % r))
        sl = slice(int(z['lo']), int(z['hi']))
        np.testing.assert_allclose(z['losses'][:5], host(lf)[:5], rtol=2e-6)            # every rank holds the full-batch losses
        np.testing.assert_array_equal(z['gdisp0'], host(gf['gdisps'][0])[sl])            # shard gradients are final, bit for bit
        assert_grad_close(z['gposes'], host(gf['gposes'])[sl], what='gposes of rank %d' % r)
        assert int(z['B_global']) == B
        np.testing.assert_allclose(z['losses_abi'][:5], host(lf)[:5], rtol=2e-6)         # C-ABI communicator, direct call
        np.testing.assert_allclose(z['losses_graph'][:5], host(lf)[:5], rtol=2e-6)       # ... and replayed inside a CUDA graph
        np.testing.assert_array_equal(z['gdisp0_abi'], host(gf['gdisps'][0])[sl])
        np.testing.assert_allclose(z['losses_peer'][:5], host(lf)[:5], rtol=2e-6)        # summed inside the epilogue kernel (peer memory)
        np.testing.assert_allclose(z['losses_peer_graph'][:5], host(lf)[:5], rtol=2e-6)
        np.testing.assert_array_equal(z['gdisp0_peer'], host(gf['gdisps'][0])[sl])
    # the in-kernel sum adds in rank order on every rank: bitwise identical results across the ranks
    z0, z1 = np.load(tmp_path / 'rank0.npz'), np.load(tmp_path / 'rank1.npz')
    np.testing.assert_array_equal(z0['losses_peer'], z1['losses_peer'])
