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
    c2, c3, c4, c5 = swin_backbone(img.float(), sd) if arch == 'swin_tiny' else backbone(img.float(), sd, arch)
    ps = fpn(c3, c4, c5, sd)
    proto = protonet(ps[0], sd)
    cls, box, coef = zip(*(head(p, sd, num_classes) for p in ps))
    cls = F.softmax(torch.cat(cls, 1), -1)
    out = (cls, torch.cat(box, 1), torch.cat(coef, 1), proto)
    if return_intermediates:
        return out, dict(c2=c2, c3=c3, c4=c4, c5=c5, p3=ps[0], p4=ps[1], p5=ps[2], p6=ps[3], p7=ps[4])
    return out


# ----------------------------------------------------------------------------- synthetic weights
def synth_state_dict(arch='res101', num_classes=81, num_ratios=3, seed=0, coef_dim=32, train=False):
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

    fpn_in = (512, 1024, 2048)
    if arch == 'swin_tiny':
        synth_swin_backbone(sd, seed)
        fpn_in = (192, 384, 768)
    else:
        conv('backbone.conv1', 64, 3, 7, False, gain=2.0); bn('backbone.bn1', 64)
    inpl = 64
    for s, nblk in enumerate(STAGES.get(arch, ())):
        planes = 64 * 2 ** s
        for b in range(nblk):
            p = f'backbone.layers.{s}.{b}'
            conv(p + '.conv1', planes, inpl, 1, False, gain=1.4); bn(p + '.bn1', planes)
            conv(p + '.conv2', planes, planes, 3, False, gain=1.4); bn(p + '.bn2', planes)
            conv(p + '.conv3', planes * 4, planes, 1, False, gain=1.0); bn(p + '.bn3', planes * 4, 0.1, 0.3)
            if b == 0:
                conv(p + '.downsample.0', planes * 4, inpl, 1, False, gain=1.0); bn(p + '.downsample.1', planes * 4, 0.5, 1.0)
            inpl = planes * 4
    for i, c in enumerate(fpn_in):
        conv(f'fpn.lat_layers.{i}', 256, c, 1, True, gain=0.45 if arch == 'swin_tiny' else 1.0)
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
    if train:
        conv('semantic_seg_conv', num_classes - 1, 256, 1, True)
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

    if arch == 'swin_tiny':
        outs = swin_backbone_emulated(img.float(), sd, q)
    else:
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


# ----------------------------------------------------------------------------- Swin-T backbone
SWIN_DEPTHS, SWIN_HEADS, SWIN_DIM, SWIN_WS = (2, 2, 6, 2), (3, 6, 12, 24), 96, 7


def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def _lin(x, sd, p, bias=True):
    return F.linear(x, sd[p + '.weight'], sd[p + '.bias'] if bias else None)


def _rel_index(ws):
    """modules/swin_transformer.py:151-160"""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _shift_mask(Hp, Wp, ws, shift):
    """modules/swin_transformer.py:369-387"""
    img_mask = torch.zeros((1, Hp, Wp, 1))
    cnt = 0
    for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = img_mask.view(1, Hp // ws, ws, Wp // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def _swin_block(x, H, W, sd, p, heads, shift, mask):
    """modules/swin_transformer.py:234-289 + :172-200"""
    B, L, C = x.shape
    ws = SWIN_WS
    h = _ln(x, sd, p + '.norm1').view(B, H, W, C)
    pad_r, pad_b = (ws - W % ws) % ws, (ws - H % ws) % ws
    h = F.pad(h, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = H + pad_b, W + pad_r
    if shift:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    win = h.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)
    B_, N = win.shape[0], ws * ws
    qkv = _lin(win, sd, p + '.attn.qkv').reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + '.attn.relative_position_bias_table'][_rel_index(ws).view(-1)].view(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(-1)
    o = _lin((attn @ v).transpose(1, 2).reshape(B_, N, C), sd, p + '.attn.proj')
    o = o.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    o = o[:, :H, :W, :].reshape(B, H * W, C)
    x = x + o
    m = _lin(F.gelu(_lin(_ln(x, sd, p + '.norm2'), sd, p + '.mlp.fc1')), sd, p + '.mlp.fc2')
    return x + m


def swin_backbone(img, sd):
    """modules/swin_transformer.py:500-518 (+ PatchEmbed :419-433, BasicLayer :361-397, PatchMerging :299-325).
    Returns the four stage outputs NCHW (stage 0 un-normalised, as the reference)."""
    B, _, H, W = img.shape
    img = F.pad(img, (0, (4 - W % 4) % 4, 0, (4 - H % 4) % 4))
    x = F.conv2d(img, sd['backbone.patch_embed.proj.weight'], sd['backbone.patch_embed.proj.bias'], stride=4)
    Wh, Ww = x.shape[2], x.shape[3]
    x = _ln(x.flatten(2).transpose(1, 2), sd, 'backbone.patch_embed.norm')
    outs = []
    for s, depth in enumerate(SWIN_DEPTHS):
        ws, shift = SWIN_WS, SWIN_WS // 2
        Hp, Wp = -(-Wh // ws) * ws, -(-Ww // ws) * ws
        mask = _shift_mask(Hp, Wp, ws, shift)
        for b in range(depth):
            x = _swin_block(x, Wh, Ww, sd, f'backbone.layers.{s}.blocks.{b}', SWIN_HEADS[s], shift if b % 2 else 0, mask)
        C = x.shape[-1]
        xo = _ln(x, sd, f'backbone.norm{s}') if s > 0 else x
        outs.append(xo.view(B, Wh, Ww, C).permute(0, 3, 1, 2).contiguous())
        if s < 3:
            p = f'backbone.layers.{s}.downsample'
            g = x.view(B, Wh, Ww, C)
            g = F.pad(g, (0, 0, 0, Ww % 2, 0, Wh % 2))
            g = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
            x = _lin(_ln(g.view(B, -1, 4 * C), sd, p + '.norm'), sd, p + '.reduction', bias=False)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs


def synth_swin_backbone(sd, seed=0):
    """Adds deterministic Swin-T backbone weights (SURVEY.md App. C key layout) to sd."""
    import numpy as np
    from . import synth
    ctr = [5000]

    def _u(shape, lo, hi):
        ctr[0] += 1
        return torch.from_numpy((lo + (hi - lo) * synth.uniform(seed, ctr[0], shape)).astype(np.float32))

    def lin(name, cout, cin, bias=True, gain=1.0):
        a = gain * (3.0 / cin) ** 0.5
        sd[name + '.weight'] = _u((cout, cin), -a, a)
        if bias:
            sd[name + '.bias'] = _u((cout,), -0.05, 0.05)

    def ln(name, c):
        sd[name + '.weight'] = _u((c,), 0.8, 1.2)
        sd[name + '.bias'] = _u((c,), -0.1, 0.1)

    a = (3.0 / 48) ** 0.5
    sd['backbone.patch_embed.proj.weight'] = _u((96, 3, 4, 4), -a, a)
    sd['backbone.patch_embed.proj.bias'] = _u((96,), -0.05, 0.05)
    ln('backbone.patch_embed.norm', 96)
    for s, depth in enumerate(SWIN_DEPTHS):
        C, nh = SWIN_DIM * 2 ** s, SWIN_HEADS[s]
        for b in range(depth):
            p = f'backbone.layers.{s}.blocks.{b}'
            ln(p + '.norm1', C); ln(p + '.norm2', C)
            sd[p + '.attn.relative_position_bias_table'] = _u((169, nh), -0.5, 0.5)
            sd[p + '.attn.relative_position_index'] = _rel_index(SWIN_WS)
            lin(p + '.attn.qkv', 3 * C, C, gain=1.2)
            lin(p + '.attn.proj', C, C, gain=0.5)
            lin(p + '.mlp.fc1', 4 * C, C, gain=1.2)
            lin(p + '.mlp.fc2', C, 4 * C, gain=0.5)
        if s < 3:
            p = f'backbone.layers.{s}.downsample'
            ln(p + '.norm', 4 * C)
            lin(p + '.reduction', 2 * C, 4 * C, bias=False)
        if s > 0:
            ln(f'backbone.norm{s}', C)
    return sd


def swin_backbone_emulated(img, sd, q):
    """Swin-T with the engine's 16-bit rounding points (q = round-to-act): every stored activation
    (LayerNorm outputs, qkv, attention output, block outputs, GELU output, merged tokens) and every
    linear weight is rounded; statistics, softmax, bias / residual adds and GELU are fp32."""
    B, _, H, W = img.shape
    img = F.pad(img, (0, (4 - W % 4) % 4, 0, (4 - H % 4) % 4))
    x = F.conv2d(img, sd['backbone.patch_embed.proj.weight'], sd['backbone.patch_embed.proj.bias'], stride=4)   # fp32 kernel
    Wh, Ww = x.shape[2], x.shape[3]
    x = q(_ln(x.flatten(2).transpose(1, 2), sd, 'backbone.patch_embed.norm'))
    lin = lambda t, p, bias=True: F.linear(t, q(sd[p + '.weight']), sd[p + '.bias'] if bias else None)
    outs = []
    for s, depth in enumerate(SWIN_DEPTHS):
        ws, shift0 = SWIN_WS, SWIN_WS // 2
        Hp, Wp = -(-Wh // ws) * ws, -(-Ww // ws) * ws
        mask = _shift_mask(Hp, Wp, ws, shift0)
        heads = SWIN_HEADS[s]
        for b in range(depth):
            p = f'backbone.layers.{s}.blocks.{b}'
            shift = shift0 if b % 2 else 0
            C = x.shape[-1]
            h = q(_ln(x, sd, p + '.norm1'))
            qkv_tok = q(lin(h, p + '.attn.qkv')).view(B, Wh, Ww, 3 * C)
            pad_r, pad_b = (ws - Ww % ws) % ws, (ws - Wh % ws) % ws
            padded = sd[p + '.attn.qkv.bias'].view(1, 1, 1, -1).expand(B, Hp, Wp, 3 * C).clone()   # pad tokens: qkv == bias (fp32)
            padded[:, :Wh, :Ww] = qkv_tok
            if shift:
                padded = torch.roll(padded, shifts=(-shift, -shift), dims=(1, 2))
            win = padded.view(B, Hp // ws, ws, Wp // ws, ws, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, 3 * C)
            B_, N = win.shape[0], ws * ws
            qkv = win.reshape(B_, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
            qq, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
            attn = qq @ k.transpose(-2, -1)
            bias = sd[p + '.attn.relative_position_bias_table'][_rel_index(ws).view(-1)].view(N, N, -1).permute(2, 0, 1)
            attn = attn + bias.unsqueeze(0)
            if shift:
                nW = mask.shape[0]
                attn = (attn.view(B_ // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
            o = q((attn.softmax(-1) @ v).transpose(1, 2).reshape(B_, N, C))
            o = o.view(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
            if shift:
                o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
            o = o[:, :Wh, :Ww, :].reshape(B, Wh * Ww, C)
            x = q(x + lin(o, p + '.attn.proj'))
            m = q(F.gelu(lin(q(_ln(x, sd, p + '.norm2')), p + '.mlp.fc1')))
            x = q(x + lin(m, p + '.mlp.fc2'))
        C = x.shape[-1]
        xo = q(_ln(x, sd, f'backbone.norm{s}')) if s > 0 else x
        outs.append(xo.view(B, Wh, Ww, C).permute(0, 3, 1, 2).contiguous())
        if s < 3:
            p = f'backbone.layers.{s}.downsample'
            g = x.view(B, Wh, Ww, C)
            g = F.pad(g, (0, 0, 0, Ww % 2, 0, Wh % 2))
            g = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1)
            x = q(lin(q(_ln(g.view(B, -1, 4 * C), sd, p + '.norm')), p + '.reduction', bias=False))
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs
