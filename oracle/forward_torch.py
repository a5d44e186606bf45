"""Plain-torch fp32 functional restatement of the reference eval forward (TEST
INFRASTRUCTURE -- see oracle/__init__.py; never imported by the product package).

Follows:
  ResNet.forward / Bottleneck.forward   modules/resnet.py:86-98, :20-40
  FPN.forward                           modules/yolact.py:73-89
  ProtoNet.forward                      modules/yolact.py:49-53 (+ permute :145)
  PredictionModule.forward              modules/yolact.py:26-31 (shared over 5 levels :149-157)
  Yolact.forward eval tail              modules/yolact.py:163-164
Works from a state_dict with the reference's key names.  The only
deliberate difference: the FPN top-down step interpolates to the lateral's size instead of
scale_factor=2 -- bit-identical whenever img_size % 32 == 0 (asserted by
tests/golden/make_golden.py against the unmodified reference) and the only way 550/400 run
at all.
"""
import torch
import torch.nn.functional as F

STAGES = {'res50': (3, 4, 6, 3), 'res101': (3, 4, 23, 3)}


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                        sd[p + '.bias'], training=False, eps=1e-5)


def _bottleneck(x, sd, p, stride):
    out = F.relu(_bn(F.conv2d(x, sd[p + '.conv1.weight']), sd, p + '.bn1'))
    out = F.relu(_bn(F.conv2d(out, sd[p + '.conv2.weight'], stride=stride, padding=1), sd, p + '.bn2'))
    out = _bn(F.conv2d(out, sd[p + '.conv3.weight']), sd, p + '.bn3')
    if (p + '.downsample.0.weight') in sd:
        x = _bn(F.conv2d(x, sd[p + '.downsample.0.weight'], stride=stride), sd, p + '.downsample.1')
    return F.relu(out + x)


def backbone(img, sd, arch):
    x = F.relu(_bn(F.conv2d(img, sd['backbone.conv1.weight'], stride=2, padding=3), sd, 'backbone.bn1'))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for s, nblk in enumerate(STAGES[arch]):
        for b in range(nblk):
            x = _bottleneck(x, sd, f'backbone.layers.{s}.{b}', 2 if (b == 0 and s > 0) else 1)
        outs.append(x)
    return outs


def fpn(c3, c4, c5, sd):
    def lat(i, x):
        return F.conv2d(x, sd[f'fpn.lat_layers.{i}.weight'], sd[f'fpn.lat_layers.{i}.bias'])

    def up(x, like):
        return F.interpolate(x, size=like.shape[2:], mode='bilinear', align_corners=False)

    def pred(i, x):
        return F.relu(F.conv2d(x, sd[f'fpn.pred_layers.{i}.0.weight'], sd[f'fpn.pred_layers.{i}.0.bias'], padding=1))

    def down(i, x):
        return F.relu(F.conv2d(x, sd[f'fpn.downsample_layers.{i}.0.weight'],
                               sd[f'fpn.downsample_layers.{i}.0.bias'], stride=2, padding=1))

    p5_1 = lat(2, c5)
    l4 = lat(1, c4)
    p4_1 = l4 + up(p5_1, l4)
    l3 = lat(0, c3)
    p3_1 = l3 + up(p4_1, l3)
    p5, p4, p3 = pred(2, p5_1), pred(1, p4_1), pred(0, p3_1)
    p6 = down(0, p5)
    p7 = down(1, p6)
    return p3, p4, p5, p6, p7


def protonet(p3, sd):
    x = p3
    for i in (0, 2, 4):
        x = F.relu(F.conv2d(x, sd[f'proto_net.proto1.{i}.weight'], sd[f'proto_net.proto1.{i}.bias'], padding=1))
    x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
    x = F.relu(F.conv2d(x, sd['proto_net.proto2.0.weight'], sd['proto_net.proto2.0.bias'], padding=1))
    x = F.relu(F.conv2d(x, sd['proto_net.proto2.2.weight'], sd['proto_net.proto2.2.bias']))
    return x.permute(0, 2, 3, 1).contiguous()


def head(x, sd, num_classes, coef_dim=32):
    B = x.size(0)
    pl = 'prediction_layers.'
    x = F.relu(F.conv2d(x, sd[pl + 'upfeature.0.weight'], sd[pl + 'upfeature.0.bias'], padding=1))
    conf = F.conv2d(x, sd[pl + 'conf_layer.weight'], sd[pl + 'conf_layer.bias'], padding=1)
    box = F.conv2d(x, sd[pl + 'bbox_layer.weight'], sd[pl + 'bbox_layer.bias'], padding=1)
    coef = torch.tanh(F.conv2d(x, sd[pl + 'coef_layer.0.weight'], sd[pl + 'coef_layer.0.bias'], padding=1))
    return (conf.permute(0, 2, 3, 1).reshape(B, -1, num_classes),
            box.permute(0, 2, 3, 1).reshape(B, -1, 4),
            coef.permute(0, 2, 3, 1).reshape(B, -1, coef_dim))


@torch.no_grad()
def forward(img, sd, arch='res101', num_classes=81, return_intermediates=False):
    """Eval forward.  img [B,3,S,S] fp32; returns (class [B,A,C] softmaxed, box [B,A,4],
    coef [B,A,32], proto [B,P,P,32])."""
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    c2, c3, c4, c5 = backbone(img.float(), sd, arch)
    ps = fpn(c3, c4, c5, sd)
    proto = protonet(ps[0], sd)
    cls, box, coef = zip(*(head(p, sd, num_classes) for p in ps))
    cls = F.softmax(torch.cat(cls, 1), -1)
    out = (cls, torch.cat(box, 1), torch.cat(coef, 1), proto)
    if return_intermediates:
        return out, dict(c2=c2, c3=c3, c4=c4, c5=c5, p3=ps[0], p4=ps[1], p5=ps[2], p6=ps[3], p7=ps[4])
    return out


# ----------------------------------------------------------------------------- synthetic weights
def synth_state_dict(arch='res101', num_classes=81, num_ratios=3, seed=0, coef_dim=32):
    """Deterministic random weights with the reference's state-dict layout:
    xavier-uniform-like conv weights (modules/yolact.py:120-125), non-trivial BN statistics
    (so BN folding is actually exercised), small non-zero biases."""
    import numpy as np
    from . import synth
    sd = {}
    ctr = [0]

    def _u(shape, lo, hi):
        ctr[0] += 1
        return torch.from_numpy((lo + (hi - lo) * synth.uniform(seed, 100 + ctr[0], shape)).astype(np.float32))

    def conv(name, cout, cin, k, bias, gain=1.0):
        fan_in, fan_out = cin * k * k, cout * k * k
        a = gain * (6.0 / (fan_in + fan_out)) ** 0.5
        sd[name + '.weight'] = _u((cout, cin, k, k), -a, a)
        if bias:
            sd[name + '.bias'] = _u((cout,), -0.05, 0.05)

    def bn(name, c, glo=0.8, ghi=1.4):
        sd[name + '.weight'] = _u((c,), glo, ghi)
        sd[name + '.bias'] = _u((c,), -0.1, 0.1)
        sd[name + '.running_mean'] = _u((c,), -0.1, 0.1)
        sd[name + '.running_var'] = _u((c,), 0.5, 1.5)
        sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    conv('backbone.conv1', 64, 3, 7, False, gain=2.0); bn('backbone.bn1', 64)
    inpl = 64
    for s, nblk in enumerate(STAGES[arch]):
        planes = 64 * 2 ** s
        for b in range(nblk):
            p = f'backbone.layers.{s}.{b}'
            conv(p + '.conv1', planes, inpl, 1, False, gain=1.4); bn(p + '.bn1', planes)
            conv(p + '.conv2', planes, planes, 3, False, gain=1.4); bn(p + '.bn2', planes)
            conv(p + '.conv3', planes * 4, planes, 1, False, gain=1.0); bn(p + '.bn3', planes * 4, 0.1, 0.3)
            if b == 0:
                conv(p + '.downsample.0', planes * 4, inpl, 1, False, gain=1.0); bn(p + '.downsample.1', planes * 4, 0.5, 1.0)
            inpl = planes * 4
    for i, c in enumerate((512, 1024, 2048)):
        conv(f'fpn.lat_layers.{i}', 256, c, 1, True)
        conv(f'fpn.pred_layers.{i}.0', 256, 256, 3, True, gain=1.3)
    for i in range(2):
        conv(f'fpn.downsample_layers.{i}.0', 256, 256, 3, True, gain=1.4)
    for i in (0, 2, 4):
        conv(f'proto_net.proto1.{i}', 256, 256, 3, True, gain=1.3)
    conv('proto_net.proto2.0', 256, 256, 3, True, gain=1.3)
    conv('proto_net.proto2.2', coef_dim, 256, 1, True, gain=1.3)
    pl = 'prediction_layers.'
    conv(pl + 'upfeature.0', 256, 256, 3, True, gain=1.3)
    conv(pl + 'bbox_layer', num_ratios * 4, 256, 3, True, gain=0.6)
    conv(pl + 'conf_layer', num_ratios * num_classes, 256, 3, True, gain=2.5)
    conv(pl + 'coef_layer.0', num_ratios * coef_dim, 256, 3, True, gain=0.8)
    return sd


# ----------------------------------------------------------------------------- 16-bit emulation
@torch.no_grad()
def forward_emulated(img, sd, arch='res101', num_classes=81, act=torch.bfloat16):
    """What a 16-bit-operand / fp32-accumulate pipeline with the engine's rounding points computes
    (TEST INFRASTRUCTURE): BN folded into fp32 weights which are then rounded to `act`; every
    stored activation rounded to `act`; accumulation, bias, residual add, ReLU, bilinear
    interpolation in fp32; head logits / proto output / softmax / tanh in fp32.  Used to separate
    "the kernels compute the 16-bit pipeline exactly" from "how far a 16-bit pipeline is from
    the fp32 reference"."""
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    q = lambda t: t.to(act).float()

    def fold(wname, bn):
        s = sd[bn + '.weight'] / torch.sqrt(sd[bn + '.running_var'] + 1e-5)
        return sd[wname] * s[:, None, None, None], sd[bn + '.bias'] - sd[bn + '.running_mean'] * s

    w, b = fold('backbone.conv1.weight', 'backbone.bn1')          # stem: im2col GEMM on 16-bit image / weights
    x = q(F.relu(F.conv2d(q(img.float()), q(w), b, stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for s, n in enumerate(STAGES[arch]):
        for bi in range(n):
            p = f'backbone.layers.{s}.{bi}'
            st = 2 if (bi == 0 and s > 0) else 1
            w1, b1 = fold(p + '.conv1.weight', p + '.bn1')
            o = q(F.relu(F.conv2d(x, q(w1), b1)))
            w2, b2 = fold(p + '.conv2.weight', p + '.bn2')
            o = q(F.relu(F.conv2d(o, q(w2), b2, stride=st, padding=1)))
            w3, b3 = fold(p + '.conv3.weight', p + '.bn3')
            o = F.conv2d(o, q(w3), b3)
            r = x
            if bi == 0:
                wd, bd = fold(p + '.downsample.0.weight', p + '.downsample.1')
                r = q(F.conv2d(x, q(wd), bd, stride=st))
            x = q(F.relu(o + r))
        outs.append(x)
    c3, c4, c5 = outs[1:]
    cv = lambda x, n, **k: F.conv2d(x, q(sd[n + '.weight']), sd[n + '.bias'], **k)
    up = lambda x, like: F.interpolate(x, size=like.shape[2:], mode='bilinear', align_corners=False)
    p5_1 = q(cv(c5, 'fpn.lat_layers.2'))
    l4 = q(cv(c4, 'fpn.lat_layers.1')); p4_1 = q(l4 + up(p5_1, l4))
    l3 = q(cv(c3, 'fpn.lat_layers.0')); p3_1 = q(l3 + up(p4_1, l3))
    p5 = q(F.relu(cv(p5_1, 'fpn.pred_layers.2.0', padding=1)))
    p4 = q(F.relu(cv(p4_1, 'fpn.pred_layers.1.0', padding=1)))
    p3 = q(F.relu(cv(p3_1, 'fpn.pred_layers.0.0', padding=1)))
    p6 = q(F.relu(cv(p5, 'fpn.downsample_layers.0.0', stride=2, padding=1)))
    p7 = q(F.relu(cv(p6, 'fpn.downsample_layers.1.0', stride=2, padding=1)))
    x = p3
    for i in (0, 2, 4):
        x = q(F.relu(cv(x, f'proto_net.proto1.{i}', padding=1)))
    x = q(F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True))
    x = q(F.relu(cv(x, 'proto_net.proto2.0', padding=1)))
    proto = F.relu(cv(x, 'proto_net.proto2.2')).permute(0, 2, 3, 1).contiguous()
    cl, bx, cf = [], [], []
    for p in (p3, p4, p5, p6, p7):
        f = q(F.relu(cv(p, 'prediction_layers.upfeature.0', padding=1)))
        B = f.size(0)
        cl.append(cv(f, 'prediction_layers.conf_layer', padding=1).permute(0, 2, 3, 1).reshape(B, -1, num_classes))
        bx.append(cv(f, 'prediction_layers.bbox_layer', padding=1).permute(0, 2, 3, 1).reshape(B, -1, 4))
        cf.append(torch.tanh(cv(f, 'prediction_layers.coef_layer.0', padding=1)).permute(0, 2, 3, 1).reshape(B, -1, 32))
    return F.softmax(torch.cat(cl, 1), -1), torch.cat(bx, 1), torch.cat(cf, 1), proto
