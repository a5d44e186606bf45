"""Multi-GPU inference plumbing: images are independent through forward and post-process, so a
global batch is sharded across ranks (one process per GPU) with no data-path collective; the
only exchange is ONE all-gather of fixed-size detection records per batch.
torch.distributed provides the process group (NCCL over NVLink on GPUs, gloo in CPU tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, init_method='env://', rank=rank, world_size=world)
    return rank, world, local


def shard_range(global_batch, rank, world):
    """Contiguous shard [lo, hi) of a global batch for `rank` (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def record_width(max_det, coef_dim):
    return 1 + max_det * (7 + coef_dim)


def pack_records(det):
    """dict from output_utils.detect_batched -> int32 [B, 1 + D*(7+K)] (floats bit-cast)."""
    B, D = det['cls'].shape
    K = det['coef'].shape[-1]
    i32 = lambda t: t.contiguous().view(torch.int32)
    return torch.cat([det['count'].view(B, 1).to(torch.int32), det['cls'].reshape(B, D), det['anchor'].reshape(B, D),
                      i32(det['score']).reshape(B, D), i32(det['box']).reshape(B, D * 4),
                      i32(det['coef']).reshape(B, D * K)], dim=1).contiguous()


def unpack_records(rec, max_det, coef_dim):
    B, D, K = rec.shape[0], max_det, coef_dim
    o = 1
    out = {'count': rec[:, 0].contiguous()}
    out['cls'] = rec[:, o:o + D].contiguous(); o += D
    out['anchor'] = rec[:, o:o + D].contiguous(); o += D
    out['score'] = rec[:, o:o + D].contiguous().view(torch.float32); o += D
    out['box'] = rec[:, o:o + 4 * D].contiguous().view(torch.float32).reshape(B, D, 4); o += 4 * D
    out['coef'] = rec[:, o:o + K * D].contiguous().view(torch.float32).reshape(B, D, K)
    return out


class _Gather:
    """Handle of an all-gather in flight: wait() orders the current stream after it, result() returns the global dict."""

    def __init__(self, out, work, world, B, D, K):
        self.out, self.work, self.dims = out, work, (world, B, D, K)

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        return self

    def result(self):
        self.wait()
        world, B, D, K = self.dims
        from .utils.output_utils import record_fields
        res = {}
        for name, o, n, shape, dt in record_fields(B, D, K)[0]:
            res[name] = self.out[:, o:o + n].contiguous().view(dt).view((world * shape[0],) + tuple(shape[1:]))
        return res


def gather_detections(det, group=None, async_op=False):
    """All-gather the detection records of every rank's shard (equal shard sizes) -- ONE NCCL call on the flat record that
    output_utils.detect_batched's kernels wrote ('_flat'; dicts from elsewhere are packed first).  Returns the dict for the
    GLOBAL batch in rank order, on every rank; with async_op=True a handle whose result() does, so that the collective overlaps
    the next batch's forward (NCCL runs it on its own stream; wait()/result() order the current stream after it)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return _Done(det) if async_op else det
    world = dist.get_world_size(group)
    B, D = det['cls'].shape
    K = det['coef'].shape[-1]
    flat = det.get('_flat')
    if flat is None:                                            # field-major per rank, as detect_batched lays it out
        from .utils.output_utils import record_views, record_numel
        v = record_views(torch.zeros(record_numel(B, D, K), dtype=torch.int32, device=det['cls'].device), B, D, K)
        for k in ('count', 'cls', 'anchor', 'score', 'box', 'coef'):
            v[k].copy_(det[k])
        flat = v['_flat']
    out = torch.empty(world * flat.numel(), dtype=torch.int32, device=flat.device)
    work = dist.all_gather_into_tensor(out, flat, group=group, async_op=True)
    h = _Gather(out.view(world, flat.numel()), work, world, B, D, K)
    return h if async_op else h.result()


class _Done:
    def __init__(self, det):
        self.det = det

    def wait(self):
        return self

    def result(self):
        return self.det
