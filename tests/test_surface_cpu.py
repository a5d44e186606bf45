"""CPU: the drop-in surface equals the reference's.  tests/golden/surface.json was minted by
tests/golden/make_golden.py from the unmodified reference: the state-dict layout (name, shape, dtype) of Yolact for
every backbone in eval and train mode, and the config attributes the hot path reads."""
import json
import os

import pytest

from yolact_minimal_b200.config import make_config
from yolact_minimal_b200.modules.yolact import Yolact

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'surface.json')))


@pytest.mark.parametrize('arch', ['res50', 'res101', 'swin_tiny'])
@pytest.mark.parametrize('mode', ['detect', 'train'])
def test_state_dict_layout_matches_reference(arch, mode):
    cfg = make_config(arch + '_coco', 544, mode=mode, train_bs=2)
    net = Yolact(cfg)
    ours = [[k, list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in net.state_dict().items()]
    ref = GOLD['state_dict'][f'{arch}/{mode}']
    assert {k: (s, d) for k, s, d in ours} == {k: (s, d) for k, s, d in ref}      # a reference checkpoint loads strictly
    assert [k for k, _, _ in ours] == [k for k, _, _ in ref]                      # and is written back in the same order


@pytest.mark.parametrize('name', ['res50_coco', 'res101_coco', 'swin_tiny_coco'])
@pytest.mark.parametrize('mode', ['detect', 'train'])
def test_config_attributes_match_reference(name, mode):
    cfg = make_config(name, 544, mode=mode, train_bs=2)
    ref = GOLD['config'][f'{name}/{mode}']
    for k, v in ref.items():
        assert hasattr(cfg, k), k
        got = getattr(cfg, k)
        assert (list(got) if isinstance(got, (list, tuple)) else got) == v, (k, got, v)


def test_default_precision_is_the_documented_one(monkeypatch):
    """`Yolact(cfg)` as eval.py / detect.py construct it (no precision attribute) must run the mode every published number and
    parity claim refers to: fp16 operands, fp32 accumulate."""
    import inspect
    from yolact_minimal_b200 import engine
    monkeypatch.delenv('YOLACT_B200_PRECISION', raising=False)
    assert engine.DEFAULT_PRECISION == 'fp16'
    assert Yolact(make_config('res50_coco', 544)).precision == 'fp16'
    assert inspect.signature(engine.Engine.sync).parameters['precision'].default == 'fp16'
