"""`Yolact(cfg)` with the reference's surface (modules/yolact.py:92-164): same constructor,
attributes (.cfg, .coef_dim, .backbone, .fpn, .proto_net, .prediction_layers, .anchors,
.semantic_seg_conv in train mode), state-dict layout, `load_weights`, and eval
`forward(img) -> (class_pred [B,A,C] softmaxed, box_pred [B,A,4], coef_pred [B,A,32],
proto_out [B,P,P,32])`.  The submodules are parameter containers; the forward pass is the
CUDA layer program in libyolact_b200.so, driven through yolact_minimal_b200.engine.Engine.

Backbones: ResNet-50/101 and Swin-T (`swin_tiny_*` configs).  In training mode forward returns the
reference's four losses (modules/yolact.py:159-161,:166-313) computed -- with their backward pass -- by the
native training engine (train_native.py -> csrc/train.cu), ResNet backbones.
"""
import math
import os

import torch
import torch.nn as nn

from .resnet import ResNet, _no_forward
from .swin_transformer import SwinTransformer
from ..engine import Engine, DEFAULT_PRECISION
from ..utils.box_utils import all_anchors


class FPN(nn.Module):
    forward = _no_forward

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.lat_layers = nn.ModuleList([nn.Conv2d(c, 256, 1) for c in in_channels])
        self.pred_layers = nn.ModuleList([nn.Sequential(nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(inplace=True))
                                          for _ in in_channels])
        self.downsample_layers = nn.ModuleList([nn.Sequential(nn.Conv2d(256, 256, 3, padding=1, stride=2), nn.ReLU(inplace=True))
                                                for _ in range(2)])


class ProtoNet(nn.Module):
    forward = _no_forward

    def __init__(self, coef_dim):
        super().__init__()
        convs = []
        for _ in range(3):
            convs += [nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(inplace=True)]
        self.proto1 = nn.Sequential(*convs)
        self.proto2 = nn.Sequential(nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(inplace=True),
                                    nn.Conv2d(256, coef_dim, 1), nn.ReLU(inplace=True))


class PredictionModule(nn.Module):
    forward = _no_forward

    def __init__(self, cfg, coef_dim=32):
        super().__init__()
        self.num_classes, self.coef_dim = cfg.num_classes, coef_dim
        r = len(cfg.aspect_ratios)
        self.upfeature = nn.Sequential(nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(inplace=True))
        self.bbox_layer = nn.Conv2d(256, r * 4, 3, padding=1)
        self.conf_layer = nn.Conv2d(256, r * cfg.num_classes, 3, padding=1)
        self.coef_layer = nn.Sequential(nn.Conv2d(256, r * coef_dim, 3, padding=1), nn.Tanh())


class Yolact(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.coef_dim = 32
        name = type(cfg).__name__
        if name.startswith('res101'):
            self.depth, blocks = 101, (3, 4, 23, 3)
        elif name.startswith('res50'):
            self.depth, blocks = 50, (3, 4, 6, 3)
        elif name.startswith('swin_tiny'):
            self.depth, blocks = 0, None                      # depth 0 selects the Swin-T layer program in the engine
        else:
            raise ValueError(f'config class {name!r} does not select a backbone (res101*/res50*/swin_tiny*)')
        if blocks is None:
            self.backbone = SwinTransformer()
            self.fpn = FPN((192, 384, 768))
        else:
            self.backbone = ResNet(blocks)
            self.fpn = FPN((512, 1024, 2048))
        self.proto_net = ProtoNet(self.coef_dim)
        self.prediction_layers = PredictionModule(cfg, self.coef_dim)
        self.anchors = all_anchors(cfg)                       # python list, like the reference (yolact.py:111-114)
        if cfg.mode == 'train':
            self.semantic_seg_conv = nn.Conv2d(256, cfg.num_classes - 1, 1)
        for m in self.modules():                              # yolact.py:120-125
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()
        self.precision = getattr(cfg, 'precision', None) or os.environ.get('YOLACT_B200_PRECISION', DEFAULT_PRECISION)
        self.max_batch = int(getattr(cfg, 'max_batch', 0) or 0)
        self._engine = None

    # ---- reference API -------------------------------------------------------------------
    def load_weights(self, weight, cuda):
        sd = torch.load(weight) if cuda else torch.load(weight, map_location='cpu')
        for k in list(sd.keys()):
            if self.cfg.mode != 'train' and k.startswith('semantic_seg_conv'):
                del sd[k]
        self.load_state_dict(sd, strict=True)
        print(f'Model loaded with {weight}.\n')
        print(f'Number of all parameters: {sum(p.numel() for p in self.parameters())}\n')

    def engine(self, batch=1):
        if self._engine is None:
            self._engine = Engine(self.depth, self.cfg.img_size, self.cfg.num_classes, len(self.cfg.aspect_ratios), self.coef_dim)
            a = self.anchors                                  # cfg.scales / cfg.aspect_ratios decide, not the engine's COCO defaults
            self._engine.set_anchors(a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a)
        self._engine.sync(self, precision=self.precision, min_batch=max(batch, self.max_batch))
        return self._engine

    def refresh_engine(self):
        """Call after modifying parameters through `.data` (EMA, weight surgery): those writes are invisible
        to the engine's change detection (engine.Engine.invalidate)."""
        if self._engine is not None:
            self._engine.invalidate()

    def forward(self, img, box_classes=None, masks_gt=None):
        if self.training:
            # native training engine (csrc/train.cu): forward, targets, losses and backward in libyolact_b200.so
            from ..train_native import training_step
            return training_step(self, img, box_classes, masks_gt)
        if not img.is_cuda:
            raise RuntimeError('yolact_minimal_b200.Yolact runs on CUDA only (no CPU fallback): move the model and input to the GPU')
        return self.engine(img.shape[0]).forward(img)
