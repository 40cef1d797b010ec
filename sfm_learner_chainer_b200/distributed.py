"""Snippet sharding across processes (one process per GPU, torch.distributed for the plumbing).

The loss path shards by snippet b with no data-path collective: every per-pixel quantity
depends only on (b, scale, source, y, x); the only coupling is that every F.mean divides by the GLOBAL
batch (base_model.py:109,111,115,184-185).  Each rank therefore runs the fused kernels on its block of
snippets with `B_global` in the descriptor -- its gradients are final -- and the five reported scalars
are partial sums that one 5-float allreduce completes.  The reference's analogue is Chainer's
MultiprocessParallelUpdater (config_utils.py:123-126), which no shipped config enables."""


def shard_range(B_global, rank, world_size):
    """Contiguous block of snippets of `rank`: [lo, hi).  Blocks differ by at most one snippet."""
    if not (0 <= rank < world_size):
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    base, rem = divmod(B_global, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_arrays(arrays, B_global, rank, world_size):
    """Slices dict(tgt, src, intrinsics, disps[list], poses, logits[list]) along the snippet axis."""
    lo, hi = shard_range(B_global, rank, world_size)
    out = {}
    for k, v in arrays.items():
        if v is None:
            out[k] = None
        elif isinstance(v, (list, tuple)):
            out[k] = [x[lo:hi] for x in v]
        else:
            out[k] = v[lo:hi]
    return out


def allreduce_loss_partials(losses, group=None, async_op=False):
    """Sums the 5 loss partials (total, pixel, smooth, exp, ssim) over ranks, in place.
    `losses`: torch tensor (CUDA with NCCL, CPU with gloo).  Returns the work handle when async."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(losses, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class LossPartialsComm(object):
    """The C-ABI communicator of the path's one collective (include/sfmloss.h: sfm_comm_* / sfm_allreduce_partials):
    an NCCL communicator owned by libsfmloss, one process per GPU.  The 128-byte NCCL id is produced on rank 0 and
    handed to the other ranks through `exchange`, a callable (bytes-or-None) -> bytes that broadcasts rank 0's
    value; the default uses the initialised torch.distributed process group (any backend).

        comm = LossPartialsComm(rank, world)            # collective: every rank calls it, on its own device
        comm.allreduce(losses_dev, stream)              # in place, asynchronous on `stream`, graph-capturable
        comm.close()                                    # after every CUDA graph that captured allreduce() is destroyed:
                                                        # such a graph holds a reference on the NCCL communicator and
                                                        # ncclCommDestroy waits for it
    """

    def __init__(self, rank, world_size, exchange=None, nccl_library=None):
        import ctypes as C
        from . import lib as L
        self._L, self._lib = L, L.load()
        self.rank, self.world_size = int(rank), int(world_size)
        if nccl_library:
            L.check(self._lib.sfm_nccl_set_library(str(nccl_library).encode()))
        buf = (C.c_char * L.SFM_NCCL_UNIQUE_ID_BYTES)()
        if self.rank == 0:
            L.check(self._lib.sfm_comm_unique_id(C.cast(buf, C.c_void_p)))
        ident = (exchange or self._torch_exchange)(bytes(buf.raw) if self.rank == 0 else None)
        if len(ident) != L.SFM_NCCL_UNIQUE_ID_BYTES:
            raise ValueError('exchange() must return the %d bytes of rank 0' % L.SFM_NCCL_UNIQUE_ID_BYTES)
        ibuf = (C.c_char * L.SFM_NCCL_UNIQUE_ID_BYTES).from_buffer_copy(ident)
        self._comm = C.c_void_p()
        L.check(self._lib.sfm_comm_create(C.cast(ibuf, C.c_void_p), self.world_size, self.rank, C.byref(self._comm)))

    @staticmethod
    def _torch_exchange(ident):
        import torch.distributed as dist
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def allreduce(self, losses, stream=None, count=5):
        """Sums the first `count` floats of the device array `losses` over the ranks, in place, on `stream` (an integer
        cudaStream_t; default: the array's current stream)."""
        import ctypes as C
        from . import device as D
        D.check_array(losses, 'losses')
        st = D.current_stream(losses) if stream is None else stream
        self._L.check(self._lib.sfm_allreduce_partials(self._comm, C.c_void_p(D.ptr(losses)), int(count), C.c_void_p(st)))

    def close(self):
        if getattr(self, '_comm', None):
            self._lib.sfm_comm_destroy(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:            # noqa: BLE001 -- interpreter shutdown
            pass


class PeerLossSum(object):
    """The loss-partial sum done INSIDE the step's epilogue kernel over NVLink peer memory (include/sfmloss.h:
    sfm_peer_* / sfm_loss_forward_backward_peer): no collective call, no extra launch.  One process per GPU on one
    node.  `gather` is a callable bytes -> list of every rank's bytes in rank order (default: torch.distributed
    all_gather_object); `barrier` a callable (default: torch.distributed.barrier).

        peer = PeerLossSum(rank, world)                 # collective: every rank calls it, on its own device
        losses, grads = op.forward_backward(..., peer=peer)    # losses = the GLOBAL sums, identical on every rank
        peer.close()                                    # collective (barrier first: a peer may still be writing)
    """

    def __init__(self, rank, world_size, gather=None, barrier=None):
        import ctypes as C
        from . import lib as L
        self._L, self._lib = L, L.load()
        self.rank, self.world_size = int(rank), int(world_size)
        self._barrier = barrier or self._torch_barrier
        self._peer = C.c_void_p()
        buf = (C.c_char * L.SFM_IPC_HANDLE_BYTES)()
        L.check(self._lib.sfm_peer_create(self.world_size, self.rank, C.byref(self._peer), C.cast(buf, C.c_void_p)))
        handles = (gather or self._torch_gather)(bytes(buf.raw))
        if len(handles) != self.world_size or any(len(h) != L.SFM_IPC_HANDLE_BYTES for h in handles):
            raise ValueError('gather() must return the %d-byte handle of every rank, in rank order' % L.SFM_IPC_HANDLE_BYTES)
        allb = (C.c_char * (L.SFM_IPC_HANDLE_BYTES * self.world_size)).from_buffer_copy(b''.join(handles))
        L.check(self._lib.sfm_peer_connect(self._peer, C.cast(allb, C.c_void_p)))
        self._barrier()                                  # every rank has mapped every slot array before the first step

    def _torch_gather(self, handle):
        import torch.distributed as dist
        if self.world_size == 1:
            return [handle]
        out = [None] * self.world_size
        dist.all_gather_object(out, handle)
        return out

    def _torch_barrier(self):
        import torch.distributed as dist
        if self.world_size > 1 and dist.is_initialized():
            dist.barrier()

    @property
    def handle(self):
        return self._peer

    def close(self):
        if getattr(self, '_peer', None):
            self._barrier()
            self._lib.sfm_peer_destroy(self._peer)
            self._peer = None


class ShardedViewSynthesisLoss(object):
    """ViewSynthesisLoss over this rank's snippet shard; losses are completed by an allreduce.

    B_global: the batch every F.mean divides by.  When omitted it is the SUM of the local batches over the group
    (one integer all-reduce at the first call), so uneven shards (shard_range gives e.g. 3 + 2) are normalised by
    the true global batch.  Every other ViewSynthesisLoss keyword (raw_disp_scales, raw_pose, edge_aware_smooth,
    n_scales) is forwarded."""

    def __init__(self, smooth_reg=0.0, exp_reg=0.0, ssim_rate=0.0, B_global=None, group=None, comm=None, peer=None, **kwargs):
        from .functions import ViewSynthesisLoss
        self.group = group
        self.peer = peer                            # PeerLossSum: the sum happens inside the epilogue kernel (no collective call)
        self.comm = comm                            # LossPartialsComm: the all-reduce then runs through the C ABI, on the loss call's stream
        self.op = ViewSynthesisLoss(smooth_reg, exp_reg, ssim_rate, B_global=B_global, **kwargs)
        self._explicit = B_global is not None
        self._b_local = None                        # local batch the cached sum was taken for

    def _resolve_global_batch(self, src):
        import torch
        import torch.distributed as dist
        b_local = int(src.shape[0])
        if b_local < 1:
            raise ValueError('empty snippet shard: the global batch must be >= the world size')
        if self._explicit:
            if self.op.B_global < b_local:
                raise ValueError('B_global=%d is smaller than the local batch %d' % (self.op.B_global, b_local))
            return
        if self._b_local == b_local:
            return                                  # same shard size as last time: the cached sum stands
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dev = src.device if dist.get_backend(self.group) == 'nccl' else 'cpu'
            t = torch.tensor([b_local], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.op.B_global = int(t.item())
        else:
            self.op.B_global = b_local
        self._b_local = b_local

    def forward_backward(self, tgt, src, intrinsics, disps, poses, logits=None, async_op=False, **kwargs):
        self._resolve_global_batch(src)
        if self.peer is not None:
            losses, grads = self.op.forward_backward(tgt, src, intrinsics, disps, poses, logits, peer=self.peer, **kwargs)
            return losses, grads, None
        losses, grads = self.op.forward_backward(tgt, src, intrinsics, disps, poses, logits, **kwargs)
        if self.comm is not None:
            self.comm.allreduce(losses)
            return losses, grads, None
        work = allreduce_loss_partials(losses, self.group, async_op=async_op)
        return losses, grads, work
