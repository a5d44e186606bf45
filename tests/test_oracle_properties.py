"""CPU: size-independent properties of the oracle (the checker the GPU parity tests lean on), on hypothesis-generated
inputs incl. the adversarial ones of SURVEY.md App. B (duplicate boxes, zero-area boxes, tied scores):
  * oracle/hard_nms.c == oracle/postprocess_np.hard_nms (two independent restatements of cython_nms.pyx:24-74)
  * greedy NMS is idempotent and its survivors are pairwise below the threshold
  * Fast-NMS output is score-sorted, truncated to max_det, never keeps a box its class would not have kept alone
  * detection-record packing is a bijection; image shards tile the batch."""
import ctypes
import os
import subprocess

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import postprocess_np as pp
from yolact_minimal_b200 import dist as ydist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_lib():
    path = os.path.join(ROOT, 'oracle', '_build', 'liboracle_hard_nms.so')
    if not os.path.exists(path):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
    lib = ctypes.CDLL(path)
    lib.oracle_hard_nms.restype = ctypes.c_int
    lib.oracle_hard_nms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]
    return lib


def _dets(rng, n, quantise, dup, degenerate):
    xy = rng.uniform(0, 500, (n, 2)).astype(np.float32)
    wh = rng.uniform(1, 120, (n, 2)).astype(np.float32)
    if degenerate:
        wh[rng.integers(0, n, max(1, n // 8))] = 0            # zero-area boxes
    d = np.concatenate([xy, xy + wh, rng.uniform(0.05, 1, (n, 1)).astype(np.float32)], 1)
    if quantise:
        d[:, 4] = np.round(d[:, 4] * 8) / 8                  # tied scores
    if dup and n > 2:
        d[rng.integers(0, n, n // 3)] = d[rng.integers(0, n, n // 3)]   # exact duplicates
    return np.ascontiguousarray(d.astype(np.float32))


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(0, 90), thr=st.sampled_from([0.0, 0.3, 0.5, 0.75, 1.0]),
       quantise=st.booleans(), dup=st.booleans(), degenerate=st.booleans())
def test_c_and_numpy_hard_nms_agree_and_are_idempotent(seed, n, thr, quantise, dup, degenerate):
    rng = np.random.default_rng(seed)
    dets = _dets(rng, n, quantise, dup, degenerate) if n else np.zeros((0, 5), np.float32)
    keep_np = pp.hard_nms(dets, thr)
    lib = _c_lib()
    flags = np.zeros(max(n, 1), np.uint8)                                    # 1 = kept, in original order
    k = lib.oracle_hard_nms(dets.ctypes.data, n, thr, flags.ctypes.data)
    assert k == len(keep_np) and np.array_equal(np.nonzero(flags[:n])[0], keep_np)
    # idempotence: survivors survive a second pass unchanged
    again = pp.hard_nms(np.ascontiguousarray(dets[keep_np]), thr)
    assert np.array_equal(again, np.arange(len(keep_np)))
    # survivors are pairwise below the threshold ("+1" pixel-area convention, ovr >= thr suppresses)
    b = dets[keep_np].astype(np.float64)
    for i in range(len(b)):
        for j in range(i + 1, len(b)):
            w = max(0.0, min(b[i, 2], b[j, 2]) - max(b[i, 0], b[j, 0]) + 1)
            h = max(0.0, min(b[i, 3], b[j, 3]) - max(b[i, 1], b[j, 1]) + 1)
            ai = (b[i, 2] - b[i, 0] + 1) * (b[i, 3] - b[i, 1] + 1)
            aj = (b[j, 2] - b[j, 0] + 1) * (b[j, 3] - b[j, 1] + 1)
            assert w * h / (ai + aj - w * h) < thr + 1e-6


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(1, 300), ncls=st.integers(1, 6), top_k=st.sampled_from([1, 5, 50, 200]),
       max_det=st.sampled_from([1, 7, 100]), iou=st.sampled_from([0.0, 0.3, 0.5, 1.0]))
def test_fast_nms_output_invariants(seed, n, ncls, top_k, max_det, iou):
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0, 0.8, (n, 2)).astype(np.float32)
    box = np.concatenate([xy, np.minimum(xy + rng.uniform(0.01, 0.4, (n, 2)).astype(np.float32), 1)], 1).astype(np.float32)
    score = rng.uniform(0.05, 1, (ncls, n)).astype(np.float32)
    anchor = np.sort(rng.choice(10 * n, n, replace=False)).astype(np.int64)
    cls, sc, bx, anc = pp.fast_nms(box, score, anchor, top_k=top_k, iou_thre=iou, max_det=max_det)
    assert len(cls) == len(anc) == len(sc) <= max_det
    assert np.all(np.diff(sc) <= 0)                                          # globally score-sorted
    pos = {a: i for i, a in enumerate(anchor)}
    for c, a, s in zip(cls, anc, sc):
        assert score[c, pos[a]] == s                                         # records are consistent
        assert (score[c] > s).sum() < top_k                                  # inside its class's top_k
    assert all(np.array_equal(b, box[pos[a]]) for b, a in zip(bx, anc))
    pairs = set(zip(cls.tolist(), anc.tolist()))
    assert len(pairs) == len(cls)                                            # no duplicates


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 2 ** 31 - 1), B=st.integers(1, 9), D=st.sampled_from([1, 5, 100]), K=st.sampled_from([8, 32]))
def test_record_packing_is_a_bijection(seed, B, D, K):
    import torch
    rng = np.random.default_rng(seed)
    count = torch.from_numpy(rng.integers(0, D + 1, B).astype(np.int32))
    cls = torch.from_numpy(rng.integers(0, 80, (B, D)).astype(np.int32))
    anc = torch.from_numpy(rng.integers(0, 19248, (B, D)).astype(np.int32))
    sc = torch.from_numpy(rng.uniform(0, 1, (B, D)).astype(np.float32))
    box = torch.from_numpy(rng.uniform(0, 1, (B, D, 4)).astype(np.float32))
    coef = torch.from_numpy(rng.uniform(-1, 1, (B, D, K)).astype(np.float32))
    det = {'count': count, 'cls': cls, 'anchor': anc, 'score': sc, 'box': box, 'coef': coef}
    rec = ydist.pack_records(det)
    assert rec.shape == (B, ydist.record_width(D, K))
    out = ydist.unpack_records(rec, D, K)
    for k in det:
        assert torch.equal(det[k], out[k]), k                                # bit-exact, floats included


@given(B=st.integers(1, 257), W=st.integers(1, 8))
def test_shards_tile_the_batch(B, W):
    spans = [ydist.shard_range(B, r, W) for r in range(W)]      # (global_batch, rank, world)
    assert spans[0][0] == 0 and spans[-1][1] == B
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [e - s for s, e in spans]
    assert max(sizes) - min(sizes) <= 1


def test_rle_oracle_round_trip_and_string_compression():
    """The RLE checker (oracle/postprocess_np.rle_counts, a restatement of pycocotools' rleEncode on Fortran-ordered masks) and the
    product's host-side ASCII compression (utils/mask_utils.py, pycocotools' rleToString / rleFrString) round-trip."""
    from yolact_minimal_b200.utils import mask_utils as mu
    from oracle import synth, postprocess_np as pp
    import numpy as np
    for seed, (h, w) in enumerate(((5, 7), (33, 20), (1, 64), (64, 1))):
        m = (synth.uniform(40 + seed, 1, (h, w)) > 0.6).astype(np.uint8)
        c = pp.rle_counts(m)
        assert int(c.sum()) == h * w and np.array_equal(pp.rle_decode(c, h, w), m)
        assert mu.rle_string_to_counts(mu.rle_counts_to_string(c)) == [int(x) for x in c]
    assert list(pp.rle_counts(np.zeros((3, 4), np.uint8))) == [12] and list(pp.rle_counts(np.ones((3, 4), np.uint8))) == [0, 12]
    # known value: counts with a long run and a decreasing tail exercise the sign-extended difference coding
    c = [0, 5, 3, 70000, 2, 1, 9]
    assert mu.rle_string_to_counts(mu.rle_counts_to_string(c)) == c
