"""oracle/train_torch.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

torch-autograd restatement of the reference's training branch (modules/yolact.py:159-161,:166-313 and
utils/box_utils.py:57-114): train-mode forward on ATen ops over the product's parameter-container modules, target
assignment and the four losses.  Round 1 shipped this file as the product's training path; round 2 replaced it with the native
engine (yolact_minimal_b200/csrc/train.cu, losses.cu) and keeps it here as the fp32 CHECKER that the native losses,
activations and parameter gradients are compared against (tests/test_train_gpu.py).  ResNet backbones only.
"""
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------
# forward (training mode): raw class logits, box regressions, tanh coefficients, prototypes, seg logits
# ----------------------------------------------------------------------------------------------------
def _conv(m, x, q):
    """A conv as the native engine computes it when q rounds to its operand type: rounded weights, (already rounded) input."""
    return F.conv2d(x, q(m.weight), m.bias, m.stride, m.padding)


def _bottleneck(blk, x, q, tap=None, name=''):
    tap = tap or (lambda n, t: t)
    out = tap(name + '.bn1.z', q(F.relu(blk.bn1(tap(name + '.conv1.y', q(_conv(blk.conv1, x, q)))))))
    out = tap(name + '.bn2.z', q(F.relu(blk.bn2(tap(name + '.conv2.y', q(_conv(blk.conv2, out, q)))))))
    out = blk.bn3(tap(name + '.conv3.y', q(_conv(blk.conv3, out, q))))
    res = x if blk.downsample is None else tap(name + '.downsample.z', q(blk.downsample[1](tap(name + '.downsample.y', q(_conv(blk.downsample[0], x, q))))))
    return tap(name + '.out', q(F.relu(out + res)))


def forward_train(net, img, taps=None, act=None, subst=None):
    """taps: optional dict that receives named intermediate activations (the native engine's tensor names), each with
    retain_grad() so that their gradients can be compared after backward.
    act: None = plain fp32 (the reference's arithmetic).  A 16-bit dtype = EMULATE the native engine's rounding points in the forward
    pass (weights and every stored activation rounded to that type, fp32 accumulation, statistics and network outputs; the cast is
    differentiable, so autograd still yields fp32 gradients): with identical ReLU masks the engine's gradients can be checked to
    ~1e-2 instead of the ~sqrt(fraction of flipped masks) that separates any 16-bit forward from an fp32 one."""
    q = (lambda t: t) if act is None else (lambda t: t.to(act).float())

    def tap(name, t):
        if subst is not None:
            # evaluate the rest of the network (and hence the whole backward pass) AT THE ENGINE'S OWN ACTIVATIONS: the value is
            # replaced by what the engine stored for this tensor, the gradient still flows into the checker's graph.  With identical
            # values every ReLU mask and batch statistic agrees, so gradients can be compared to ~1e-2 instead of the
            # sqrt(fraction of flipped masks) that separates two independently rounded 16-bit forwards.
            v = subst(name)
            if v is not None:
                t = v.to(t.dtype) + (t - t.detach())
        if taps is not None:
            if t.requires_grad:
                t.retain_grad()
            taps[name] = t
        return t
    bb = net.backbone
    if not hasattr(bb, 'conv1'):
        raise NotImplementedError('training forward is implemented for the ResNet backbones only')
    x = tap('pool', F.max_pool2d(tap('stem.z', q(F.relu(bb.bn1(tap('stem.y', q(_conv(bb.conv1, q(img), q))))))), kernel_size=3, stride=2, padding=1))
    feats = []
    for si, stage in enumerate(bb.layers):
        for i, blk in enumerate(stage):
            if i > 0 and blk.downsample is not None:            # container quirk: only block 0 owns the shortcut
                raise RuntimeError('unexpected downsample')
            x = _bottleneck(blk, x, q, tap, f'backbone.layers.{si}.{i}')
        feats.append(tap('c%d' % (len(feats) + 2), x))
    c3, c4, c5 = feats[1:]
    fpn = net.fpn
    up = lambda t, like: F.interpolate(t, size=like.shape[2:], mode='bilinear', align_corners=False)
    p5_1 = q(_conv(fpn.lat_layers[2], c5, q))
    l4 = q(_conv(fpn.lat_layers[1], c4, q))
    p4_1 = q(l4 + up(p5_1, l4))
    l3 = q(_conv(fpn.lat_layers[0], c3, q))
    p3_1 = q(l3 + up(p4_1, l3))
    tap('p5_1', p5_1); tap('p4_1', p4_1); tap('p3_1', p3_1)
    cr = lambda seq, t: q(F.relu(_conv(seq[0], t, q)))          # conv + ReLU blocks (nn.Sequential(conv, ReLU))
    p5, p4, p3 = tap('p5', cr(fpn.pred_layers[2], p5_1)), tap('p4', cr(fpn.pred_layers[1], p4_1)), tap('p3', cr(fpn.pred_layers[0], p3_1))
    p6 = tap('p6', cr(fpn.downsample_layers[0], p5))
    p7 = tap('p7', cr(fpn.downsample_layers[1], p6))
    levels = (p3, p4, p5, p6, p7)

    pn = net.proto_net
    t = p3
    for i in (0, 2, 4):
        t = tap(f'proto1.{i}', q(F.relu(_conv(pn.proto1[i], t, q))))
    t = tap('proto.up', q(F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=True)))
    t = tap('proto2.0', q(F.relu(_conv(pn.proto2[0], t, q))))
    proto = F.relu(_conv(pn.proto2[2], t, q)).permute(0, 2, 3, 1).contiguous()

    pl, B = net.prediction_layers, img.shape[0]
    cls, box, coef = [], [], []
    for li, lv in enumerate(levels):
        f = tap(f'head.f{li}', q(F.relu(_conv(pl.upfeature[0], lv, q))))
        cls.append(_conv(pl.conf_layer, f, q).permute(0, 2, 3, 1).reshape(B, -1, pl.num_classes))
        box.append(_conv(pl.bbox_layer, f, q).permute(0, 2, 3, 1).reshape(B, -1, 4))
        coef.append(torch.tanh(_conv(pl.coef_layer[0], f, q)).permute(0, 2, 3, 1).reshape(B, -1, pl.coef_dim))
    seg = _conv(net.semantic_seg_conv, p3, q)
    # the five network outputs (substitutable too: identical logits -> identical OHEM / target decisions in the losses)
    return (tap('out.cls', torch.cat(cls, 1)), tap('out.box', torch.cat(box, 1)), tap('out.coef', torch.cat(coef, 1)), tap('out.proto', proto),
            tap('out.seg', seg))


# ----------------------------------------------------------------------------------------------------
# target assignment (utils/box_utils.py:57-114)
# ----------------------------------------------------------------------------------------------------
def pairwise_iou(a, b):
    """[n,4] x [m,4] corner boxes -> [n,m] (utils/box_utils.py:8-37, same operation order)."""
    hi = torch.min(a[:, None, 2:], b[None, :, 2:])
    lo = torch.max(a[:, None, :2], b[None, :, :2])
    wh = torch.clamp(hi - lo, min=0)
    inter = wh[..., 0] * wh[..., 1]
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))[:, None]
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[None, :]
    return inter / (area_a + area_b - inter)


def encode_offsets(matched, anchors):
    """SSD encoding with variances 0.1 / 0.2 (utils/box_utils.py:104-114)."""
    cxcy = (matched[:, :2] + matched[:, 2:]) / 2 - anchors[:, :2]
    cxcy = cxcy / (0.1 * anchors[:, 2:])
    wh = torch.log((matched[:, 2:] - matched[:, :2]) / anchors[:, 2:]) / 0.2
    return torch.cat([cxcy, wh], 1)


def assign_targets(cfg, box_gt, anchors, class_gt):
    """Per image: every anchor takes its best-IoU ground truth; each ground truth additionally claims
    its best anchor (IoU forced to 2); labels: >0 foreground, 0 background (< neg thr), -1 neutral."""
    corners = torch.cat((anchors[:, :2] - anchors[:, 2:] / 2, anchors[:, :2] + anchors[:, 2:] / 2), 1)
    iou = pairwise_iou(box_gt, corners)                           # [num_gt, A]
    best_anchor_of_gt = iou.max(1)[1]
    best_iou, best_gt = iou.max(0)
    best_iou.index_fill_(0, best_anchor_of_gt, 2)
    for j in range(best_anchor_of_gt.size(0)):                    # sequential: a later gt wins a shared anchor
        best_gt[best_anchor_of_gt[j]] = j
    matched = box_gt[best_gt]
    labels = class_gt[best_gt] + 1
    labels[best_iou < cfg.pos_iou_thre] = -1
    labels[best_iou < cfg.neg_iou_thre] = 0
    return encode_offsets(matched, anchors), labels, matched, best_gt


# ----------------------------------------------------------------------------------------------------
# the four losses (modules/yolact.py:205-313)
# ----------------------------------------------------------------------------------------------------
def _crop(masks, boxes, padding=1):
    """utils/box_utils.py:147-168 on [h,w,n] masks."""
    h, w, n = masks.shape

    def span(a, b, size):
        a, b = a * size, b * size
        lo = torch.clamp(torch.min(a, b) - padding, min=0)
        hi = torch.clamp(torch.max(a, b) + padding, max=size)
        return lo, hi
    x1, x2 = span(boxes[:, 0], boxes[:, 2], w)
    y1, y2 = span(boxes[:, 1], boxes[:, 3], h)
    cols = torch.arange(w, device=masks.device, dtype=x1.dtype).view(1, -1, 1)
    rows = torch.arange(h, device=masks.device, dtype=x1.dtype).view(-1, 1, 1)
    inside = (cols >= x1.view(1, 1, -1)) & (cols < x2.view(1, 1, -1)) & (rows >= y1.view(1, 1, -1)) & (rows < y2.view(1, 1, -1))
    return masks * inside.float()


def category_loss(cfg, class_p, labels, pos, neg_pos_ratio=3):
    """Cross entropy over positives + OHEM-mined negatives (3:1), summed, / #positives."""
    flat = class_p.reshape(-1, cfg.num_classes)
    mx = flat.max()
    hardness = torch.log(torch.sum(torch.exp(flat - mx), 1)) + mx - flat[:, 0]
    hardness = hardness.reshape(class_p.size(0), -1)
    hardness[pos] = 0
    hardness[labels < 0] = 0
    rank = hardness.sort(1, descending=True)[1].sort(1)[1]
    num_pos = pos.long().sum(1, keepdim=True)
    num_neg = torch.clamp(neg_pos_ratio * num_pos, max=pos.size(1) - 1)
    neg = rank < num_neg.expand_as(rank)
    neg[pos] = 0
    neg[labels < 0] = 0
    chosen = pos + neg
    return cfg.conf_alpha * F.cross_entropy(class_p[chosen].reshape(-1, cfg.num_classes), labels[chosen], reduction='sum') / num_pos.sum()


def box_loss(cfg, box_p, offsets, pos):
    return cfg.bbox_alpha * F.smooth_l1_loss(box_p[pos, :], offsets[pos, :], reduction='sum') / pos.sum()


def mask_loss(cfg, pos, best_gt, coef_p, proto_p, mask_gt, matched):
    ph, pw = proto_p.shape[1:3]
    total = 0
    for i in range(coef_p.size(0)):
        gt = F.interpolate(mask_gt[i].unsqueeze(0), (ph, pw), mode='bilinear', align_corners=False).squeeze(0)
        gt = gt.permute(1, 2, 0).contiguous().gt(0.5).float()
        sel = pos[i]
        idx, boxes, coef = best_gt[i][sel], matched[i][sel], coef_p[i][sel]
        if idx.size(0) == 0:
            continue
        n_all = coef.size(0)
        if n_all > cfg.masks_to_train:                            # random subset, re-weighted below
            keep = torch.randperm(n_all)[:cfg.masks_to_train]
            idx, boxes, coef = idx[keep], boxes[keep], coef[keep]
        n = coef.size(0)
        pred = _crop(torch.sigmoid(proto_p[i] @ coef.t()), boxes)
        bce = F.binary_cross_entropy(torch.clamp(pred, 0, 1), gt[:, :, idx], reduction='none')
        area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
        per_obj = bce.sum(dim=(0, 1)) / area
        if n_all > n:
            per_obj = per_obj * (n_all / n)
        total = total + torch.sum(per_obj)
    return cfg.mask_alpha * total / ph / pw / pos.sum()


def semantic_loss(cfg, seg_p, mask_gt, class_gt):
    B, _, mh, mw = seg_p.size()
    total = 0
    for i in range(B):
        gt = F.interpolate(mask_gt[i].unsqueeze(0), (mh, mw), mode='bilinear', align_corners=False).squeeze(0).gt(0.5).float()
        target = torch.zeros_like(seg_p[i], requires_grad=False)
        for j in range(gt.size(0)):
            target[class_gt[i][j]] = torch.max(target[class_gt[i][j]], gt[j])
        total = total + F.binary_cross_entropy_with_logits(seg_p[i], target, reduction='sum')
    return cfg.semantic_alpha * total / mh / mw / B


def compute_loss(net, class_p, box_p, coef_p, proto_p, seg_p, box_classes, masks_gt):
    cfg, dev = net.cfg, class_p.device
    if isinstance(net.anchors, list):
        net.anchors = torch.tensor(net.anchors, device=dev).reshape(-1, 4)
    anchors = net.anchors.to(dev)
    B, A = box_p.size(0), anchors.shape[0]
    offsets = torch.zeros((B, A, 4), dtype=torch.float32, device=dev)
    labels = torch.zeros((B, A), dtype=torch.int64, device=dev)
    matched = torch.zeros((B, A, 4), dtype=torch.float32, device=dev)
    best_gt = torch.zeros((B, A), dtype=torch.int64, device=dev)
    class_gt = []
    with torch.no_grad():
        for i in range(B):
            class_gt.append(box_classes[i][:, -1].long())
            offsets[i], labels[i], matched[i], best_gt[i] = assign_targets(cfg, box_classes[i][:, :-1], anchors, class_gt[i])
    pos = labels > 0
    return (category_loss(cfg, class_p, labels, pos), box_loss(cfg, box_p, offsets, pos),
            mask_loss(cfg, pos, best_gt, coef_p, proto_p, masks_gt, matched), semantic_loss(cfg, seg_p, masks_gt, class_gt))


def training_step_forward(net, img, box_classes, masks_gt, taps=None, act=None, subst=None):
    """Yolact.forward in training mode: the reference's 4-tuple of losses (act / subst: see forward_train)."""
    outs = forward_train(net, img, taps, act, subst)
    return compute_loss(net, *outs, box_classes, masks_gt)
