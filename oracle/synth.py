"""Deterministic synthetic inputs (TEST INFRASTRUCTURE).

A counter-based generator (splitmix64 finaliser over seed/stream/index) so that the golden
generator (build container) and the tests (GPU box) produce bit-identical inputs without
depending on numpy/torch RNG stream stability.  Everything is computed in uint64/float64 and
rounded to float32 once.
"""
import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    with np.errstate(over='ignore'):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def bits(seed, stream, n):
    """n uint64 words for (seed, stream)."""
    with np.errstate(over='ignore'):
        base = _mix(np.uint64(seed) * _G + np.uint64(stream) * np.uint64(0xD1342543DE82EF95) + _G)
        idx = np.arange(1, n + 1, dtype=np.uint64)
        return _mix(base + idx * _G)


def uniform(seed, stream, shape):
    """float64 U[0,1) with 53 random bits."""
    n = int(np.prod(shape))
    u = (bits(seed, stream, n) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return u.reshape(shape)


def normal(seed, stream, shape):
    """float64 N(0,1) via Box-Muller on two uniform streams."""
    n = int(np.prod(shape))
    u1 = uniform(seed, 2 * stream + 1000003, (n,))
    u2 = uniform(seed, 2 * stream + 1000004, (n,))
    r = np.sqrt(-2.0 * np.log1p(-u1))          # 1-u1 in (0,1]
    return (r * np.cos(2.0 * np.pi * u2)).reshape(shape)


def softmax_rows(x):
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def head_outputs(seed, A, num_classes=81, regime='stress', coef_dim=32):
    """Synthetic post-softmax class scores / box regressions / mask coefficients for one image
    (SURVEY.md section 8(d) 'Fast-NMS synthetic inputs').  Returns float32 arrays
    cls [A,C], box [A,4], coef [A,coef_dim]."""
    z = normal(seed, 1, (A, num_classes))
    if regime == 'stress':            # ~100% of anchors pass the 0.05 score filter
        logits = 3.0 * z
    elif regime == 'realistic':       # ~8% pass (1570 of 18525 at S=544)
        logits = 1.6 * z
        logits[:, 0] += 9.0
    elif regime == 'sparse':          # ~0.4% pass (67 of 18525 at S=544)
        logits = 1.0 * z
        logits[:, 0] += 8.0
    else:
        raise ValueError(regime)
    cls = softmax_rows(logits).astype(np.float32)
    box = (0.5 * normal(seed, 2, (A, 4))).astype(np.float32)
    coef = np.tanh(normal(seed, 3, (A, coef_dim))).astype(np.float32)
    return cls, box, coef


def proto(seed, P, coef_dim=32):
    return np.maximum(normal(seed, 4, (P, P, coef_dim)), 0.0).astype(np.float32)


def image_batch(seed, B, S):
    """Mean/std-normalised RGB images are ~N(0,1) (utils/augmentations.py:212-216)."""
    return normal(seed, 5, (B, 3, S, S)).astype(np.float32)


def train_targets(seed, B, S, n=3, num_classes=80):
    """Synthetic training targets: per image n boxes [x1,y1,x2,y2,label]
    in [0,1] and filled-rectangle masks [n,S,S] (float32)."""
    targets, masks = [], []
    for b in range(B):
        xy = uniform(seed, 30 + b, (n, 2)) * 0.5
        wh = uniform(seed, 60 + b, (n, 2)) * 0.4 + 0.1
        lab = np.floor(uniform(seed, 90 + b, (n, 1)) * num_classes)
        t = np.concatenate([xy, xy + wh, lab], 1).astype(np.float32)
        m = np.zeros((n, S, S), np.float32)
        for j in range(n):
            x1, y1, x2, y2 = (t[j, :4] * S).astype(int)
            m[j, y1:y2 + 1, x1:x2 + 1] = 1.0
        targets.append(t); masks.append(m)
    return targets, masks
