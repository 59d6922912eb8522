import hashlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def param_digest(module):
    """sha1 over state_dict keys+bytes, identical to oracle/make_golden.py:param_digest"""
    h = hashlib.sha1()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def oracle_params(model, dt=False):
    """FieldParams (CPU) from a sunerf_b200 / reference-layout model."""
    from oracle import sunerf_oracle as orc
    ws = [model.in_layer[1].weight] + [l.weight for l in model.layers] + [model.out_layer.weight]
    bs = [model.in_layer[1].bias] + [l.bias for l in model.layers] + [model.out_layer.bias]
    la = vc = None
    if dt:
        la = torch.stack([model.log_absortpion[str(c)] for c in orc.AIA_CHANNELS]).detach().cpu().clone()
        vc = model.volumetric_constant.detach().cpu().clone()
    return orc.FieldParams([w.detach().cpu().clone() for w in ws], [b.detach().cpu().clone() for b in bs], la, vc)


def rel_err(a, b, floor=0.0):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs() / (b.abs() + floor)).max().item()


def t(x, device='cuda'):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device)
