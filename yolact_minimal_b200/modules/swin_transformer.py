"""Parameter container for the Swin-T backbone with the reference's state-dict layout
(modules/swin_transformer.py:131-518; key list in SURVEY.md App. C).  Weights only -- patch
embedding, LayerNorms, shifted-window attention, MLPs and patch merging run inside
libyolact_b200.so (csrc/swin.cu + the conv kernels for every linear layer).
`init_backbone(path)` keeps the reference's behaviour (:486-498): re-initialise Linear / LayerNorm,
then a NON-strict load of the pretrained file."""
import torch
import torch.nn as nn

from .resnet import _no_forward


def _rel_index(ws):
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing='ij')).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


class WindowAttention(nn.Module):
    forward = _no_forward

    def __init__(self, dim, window_size, num_heads):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        self.register_buffer('relative_position_index', _rel_index(window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)


class Mlp(nn.Module):
    forward = _no_forward

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class SwinTransformerBlock(nn.Module):
    forward = _no_forward

    def __init__(self, dim, num_heads, window_size, shift_size):
        super().__init__()
        self.shift_size = shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, 4 * dim)


class PatchMerging(nn.Module):
    forward = _no_forward

    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class BasicLayer(nn.Module):
    forward = _no_forward

    def __init__(self, dim, depth, num_heads, window_size, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2)
                                     for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None


class PatchEmbed(nn.Module):
    forward = _no_forward

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)


class SwinTransformer(nn.Module):
    forward = _no_forward

    def __init__(self, embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7):
        super().__init__()
        self.patch_embed = PatchEmbed(embed_dim=embed_dim)
        self.layers = nn.ModuleList([BasicLayer(embed_dim * 2 ** i, depths[i], num_heads[i], window_size, i < len(depths) - 1)
                                     for i in range(len(depths))])
        self.num_features = [embed_dim * 2 ** i for i in range(len(depths))]
        for i in (1, 2, 3):
            self.add_module(f'norm{i}', nn.LayerNorm(self.num_features[i]))

    def init_backbone(self, weight):
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        self.load_state_dict(torch.load(weight), strict=False)
        print(f'\nBackbone is initiated with {weight}.\n')
