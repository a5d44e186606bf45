"""CPU (gloo, world_size 2): the sharding / detection-record all-gather logic of the N>1 path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolact_minimal_b200 import dist as ydist


def _fake_det(B, D, K, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(count=torch.randint(0, D + 1, (B,), generator=g, dtype=torch.int32),
                cls=torch.randint(0, 80, (B, D), generator=g, dtype=torch.int32),
                anchor=torch.randint(0, 19248, (B, D), generator=g, dtype=torch.int32),
                score=torch.rand(B, D, generator=g), box=torch.rand(B, D, 4, generator=g), coef=torch.randn(B, D, K, generator=g))


def test_shard_range_covers_batch():
    for gb in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [ydist.shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    det = _fake_det(5, 100, 32, 3)
    rec = ydist.pack_records(det)
    assert rec.shape == (5, ydist.record_width(100, 32)) and rec.dtype == torch.int32
    back = ydist.unpack_records(rec, 100, 32)
    for k in det:
        assert torch.equal(det[k], back[k]), k          # bit-exact through the int32 view


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    ydist.init_from_env(backend='gloo')
    det = _fake_det(4, 100, 32, 10 + rank)
    full = ydist.gather_detections(det)
    # the product's own layout: typed views of ONE flat record (what the post-process kernels write), gathered asynchronously
    from yolact_minimal_b200.utils.output_utils import record_views, record_numel
    views = record_views(torch.empty(record_numel(4, 100, 32), dtype=torch.int32), 4, 100, 32)
    for k, v in det.items():
        views[k].copy_(v)
    h = ydist.gather_detections(views, async_op=True)
    again = h.wait().result()
    assert all(torch.equal(full[k], again[k]) for k in full)
    q.put((rank, {k: v.numpy() for k, v in full.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_detections_world2():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    expect = {k: np.concatenate([_fake_det(4, 100, 32, 10)[k].numpy(), _fake_det(4, 100, 32, 11)[k].numpy()]) for k in got[0]}
    for r in (0, 1):
        for k in expect:
            assert np.array_equal(got[r][k], expect[k]), (r, k)     # gathered == single-process result, on every rank
