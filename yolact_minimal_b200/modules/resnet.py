"""Parameter container for the ResNet-50/101 backbone with the reference's state-dict layout
(modules/resnet.py:5-104; key list in SURVEY.md App. C).  These modules hold weights only --
the convolutions themselves run inside libyolact_b200.so (csrc/net.cu builds the layer program
from the same names).  `init_backbone(path)` keeps the reference's strict-load behaviour
(modules/resnet.py:100-104)."""
import torch
import torch.nn as nn


def _no_forward(self, *a, **k):
    raise RuntimeError('parameter container: the forward pass runs in the CUDA engine (yolact_minimal_b200.engine)')


class Bottleneck(nn.Module):
    expansion = 4
    forward = _no_forward

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)   # stride lives on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    forward = _no_forward

    def __init__(self, layers):
        super().__init__()
        self.depth_blocks = tuple(layers)
        self.layers = nn.ModuleList()           # registered before the stem, as in the reference (modules/resnet.py:49-55):
        self.channels = []                      # state_dict() / parameters() enumerate in the same order
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for stage, nblk in enumerate(layers):
            planes, stride = 64 * 2 ** stage, (1 if stage == 0 else 2)
            ds = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))
            blocks = [Bottleneck(inplanes, planes, stride, ds)]
            inplanes = planes * 4
            blocks += [Bottleneck(inplanes, planes) for _ in range(1, nblk)]
            self.layers.append(nn.Sequential(*blocks))
            self.channels.append(inplanes)

    def init_backbone(self, path):
        self.load_state_dict(torch.load(path), strict=True)
        print(f'\nBackbone is initiated with {path}.\n')
