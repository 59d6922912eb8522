"""Checkpoint interop: trained reference models render through this package.

  * Lightning `.ckpt` (trainer.save_checkpoint, sunerf/run_emission.py:76): a dict whose 'state_dict' holds the
    LightningModule's tensors; the rendering module's are prefixed `rendering.` (sunerf/model/sunerf.py:16-27).
  * `save_state.snf` (save_state, sunerf/model/sunerf.py:62-74): a dict with the pickled rendering module under
    'rendering' plus data_config / Rs_per_ds / seconds_per_dt / ref_time (read by evaluation/loader.py:23-48).

The `.snf` pickle references the reference's classes by module path (`sunerf.rendering.emission...`).  They are
resolved to light-weight stand-ins (plain nn.Modules that only carry the pickled state), so neither the reference
package nor its third-party imports are needed; a module of THIS package is then built with the hyper-parameters read
off the pickled tensors and filled with `load_state_dict`.
"""
from __future__ import annotations

import pickle
from typing import Any, Dict, Optional

import torch
from torch import nn

from . import ops
from ._lib import SnfError
from .model import NeRF, NeRF_DT, SimpleStar
from .rendering import DensityTemperatureRadiativeTransfer, EmissionRadiativeTransfer, SuNeRFRendering


# ------------------------------------------------------------------------------------------ Lightning .ckpt
def rendering_state_dict(ckpt: Dict[str, Any], prefix: str = 'rendering.') -> Dict[str, torch.Tensor]:
    """The rendering module's tensors out of a Lightning checkpoint dict (or a bare state_dict)."""
    sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
    out = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    if not out:
        raise SnfError(f"no '{prefix}*' tensors in the checkpoint (keys: {list(sd)[:4]} ...)")
    return out


def load_lightning_checkpoint(rendering: SuNeRFRendering, path_or_dict, strict: bool = True, map_location='cpu') -> SuNeRFRendering:
    """Fill `rendering` from a Lightning `.ckpt` written by the reference's training scripts."""
    ckpt = torch.load(path_or_dict, map_location=map_location, weights_only=False) if isinstance(path_or_dict, str) else path_or_dict
    rendering.load_state_dict(rendering_state_dict(ckpt), strict=strict)
    return rendering


# ------------------------------------------------------------------------------------------ save_state.snf
class _Stub(nn.Module):
    """Stands in for a reference class while unpickling: nn.Module.__setstate__ restores the pickled __dict__."""
    _ref_name = ''


_STUBS: Dict[str, type] = {}


def _stub_for(module: str, name: str) -> type:
    key = f'{module}.{name}'
    if key not in _STUBS:
        _STUBS[key] = type(name, (_Stub,), {'_ref_name': key})
    return _STUBS[key]


class _Opaque:
    """Stands in for an object of a third-party package that is not installed (e.g. the xitorch interpolators the
    reference's density-temperature module keeps in `self.response`): keeps whatever state the pickle carries."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {'_state': state})


_OPAQUE: Dict[str, type] = {}


# What a reference save_state.snf legitimately pickles: tensors / storages / nn modules (torch), containers
# (collections), numpy arrays and scalars of the data config, datetimes, and a handful of harmless builtins.  Every other
# global the pickle names - whether its package is installed here or not - becomes an inert _Opaque stand-in: it is
# never imported and none of its code runs.  (A pickle is still a program: load files from sources you trust.)
_ALLOWED_ROOTS = ('torch', 'collections', 'numpy', 'datetime', '_codecs', 'copyreg', 'sunerf_b200')
_ALLOWED_BUILTINS = {'set', 'frozenset', 'slice', 'complex', 'dict', 'list', 'tuple', 'int', 'float', 'str', 'bool', 'bytes',
                     'bytearray', 'range', 'object', 'getattr'}


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == 'sunerf' or module.startswith('sunerf.'):
            return _stub_for(module, name)
        root = module.split('.')[0]
        if root in _ALLOWED_ROOTS or (module in ('builtins', '__builtin__') and name in _ALLOWED_BUILTINS and name != 'getattr'):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                pass
        key = f'{module}.{name}'
        if key not in _OPAQUE:
            _OPAQUE[key] = type(name, (_Opaque,), {'_ref_name': key})
        return _OPAQUE[key]


class _RefPickle:
    """`pickle_module` for torch.load: the stock pickle with reference classes mapped to stubs."""
    __name__ = 'sunerf_b200_ref_pickle'
    Unpickler = _RefUnpickler
    load = staticmethod(lambda f, **kw: _RefUnpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dump, dumps, Pickler = staticmethod(pickle.dump), staticmethod(pickle.dumps), pickle.Pickler
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL


def _model_class_and_config(stub_model: nn.Module):
    name = type(stub_model).__name__
    sd = stub_model.state_dict()
    if name == 'SimpleStar':
        return SimpleStar, {}
    w_in = sd['in_layer.1.weight']
    n_hidden = 1 + len([k for k in sd if k.startswith('layers.') and k.endswith('.weight')])
    cfg = {'d_filter': int(w_in.shape[0]), 'n_layers': n_hidden}
    if name == 'NeRF_DT':
        return NeRF_DT, cfg
    if name in ('NeRF', 'EmissionModel'):
        return NeRF, cfg
    raise SnfError(f'unknown field-network class in the pickle: {name}')


def _tensors_in(obj, depth: int = 0):
    """every tensor reachable through the attributes of an (opaque) unpickled object"""
    if isinstance(obj, torch.Tensor):
        yield obj
    elif depth < 4:
        vals = obj.values() if isinstance(obj, dict) else (obj if isinstance(obj, (list, tuple)) else
                                                           getattr(obj, '__dict__', {}).values())
        for v in vals:
            yield from _tensors_in(v, depth + 1)


def _adopt_pickled_response(stub: nn.Module, rend: DensityTemperatureRadiativeTransfer) -> None:
    """The reference does not store `aia_exp_time`; what it pickles are the interpolators of `self.response`
    (density_temperature.py:144-146), whose y tables are float32(TRESP * aia_exp_time).  A model trained with another
    exposure time must render with ITS tables, not with the shipped 2.9 s ones: take them from the pickle when they can
    be found there (7 channels x 101 values on the shipped logT grid), otherwise say that the shipped table is used."""
    resp = getattr(stub, 'response', None)
    if not isinstance(resp, dict):
        return
    tx, ty = rend._table_x, rend._table_y.clone()
    found = 0
    for i, ch in enumerate(ops.AIA_CHANNELS):
        interp = resp.get(ch, resp.get(str(ch), resp.get(float(ch))))
        cands = [t.detach().float().reshape(-1) for t in _tensors_in(interp) if t.numel() == tx.numel()]
        ys = [t for t in cands if not torch.equal(t, tx) and float(t.abs().max()) < 1.0]     # responses are ~1e-24, logT is 4..9
        if len(ys) == 1:
            ty[i] = ys[0]
            found += 1
    if found == len(ops.AIA_CHANNELS):
        rend._table_y.copy_(ty)
    elif found:
        raise SnfError('only some of the pickled AIA response tables could be recovered: refusing to mix them with the shipped ones')
    else:
        import warnings
        warnings.warn('the pickled response interpolators carry no recoverable tables (xitorch layout unknown): rendering with '
                      'the shipped aia_exp_time = 2.9 s response; pass aia_exp_time explicitly if the model was trained otherwise')


def rebuild_rendering(stub: nn.Module, precision: Optional[str] = None) -> SuNeRFRendering:
    """A rendering module of this package with the structure and weights of an unpickled reference module."""
    name = type(stub).__name__
    Rs = float(stub.Rs_per_ds)
    smp, hs = stub.sampler, stub.sampler_hierarchical
    kinds = {'StratifiedSampler': 'stratified', 'SphericalSampler': 'spherical'}
    if type(smp).__name__ not in kinds:
        raise SnfError(f'unknown sampler class in the pickle: {type(smp).__name__}')
    sampling_config = {'type': kinds[type(smp).__name__], 'distance': float(smp.distance) * Rs,
                       'n_samples': int(smp.t_vals.shape[-1]), 'perturb': bool(smp.perturb)}
    hier_config = {'type': 'hierarchical', 'n_samples': int(hs.n_samples), 'perturb': bool(hs.perturb)}
    model_cls, model_cfg = _model_class_and_config(stub.fine_model)
    if precision is not None and model_cls is not SimpleStar:
        model_cfg['precision'] = precision
    if name == 'EmissionRadiativeTransfer':
        model_cfg.pop('d_input', None)
        rend = EmissionRadiativeTransfer(Rs_per_ds=Rs, sampling_config=sampling_config,
                                         hierarchical_sampling_config=hier_config, model_config=model_cfg)
    elif name == 'DensityTemperatureRadiativeTransfer':
        rend = DensityTemperatureRadiativeTransfer(Rs_per_ds=Rs, sampling_config=sampling_config,
                                                   hierarchical_sampling_config=hier_config, model=model_cls,
                                                   model_config=model_cfg,
                                                   pixel_intensity_factor=float(getattr(stub, 'pixel_intensity_factor', 1e10)))
        _adopt_pickled_response(stub, rend)
    else:
        raise SnfError(f'unknown rendering class in the pickle: {name}')
    sd = {k: v for k, v in stub.state_dict().items() if k in rend.state_dict()}
    missing = [k for k in rend.state_dict() if k not in sd and not k.startswith('_')]
    if missing:
        raise SnfError(f'the pickled module lacks tensors this package needs: {missing[:5]}')
    rend.load_state_dict(sd, strict=False)
    return rend


def load_save_state(path: str, precision: Optional[str] = None, map_location='cpu') -> Dict[str, Any]:
    """Read a `save_state.snf` written by the reference (or by `save_state` below).  Returns the same dict with
    'rendering' replaced by a sunerf_b200 module carrying the trained weights."""
    state = torch.load(path, map_location=map_location, pickle_module=_RefPickle, weights_only=False)
    r = state['rendering']
    state['rendering'] = rebuild_rendering(r, precision) if isinstance(r, _Stub) else r
    return state


def save_state(rendering: SuNeRFRendering, path: str, data_config=None, Rs_per_ds=None, seconds_per_dt=None, ref_time=None) -> None:
    """Same layout as the reference's save_state (sunerf/model/sunerf.py:62-74); the module is pickled by this
    package's class path (INTEGRATION.md shows the alias that makes it loadable by the reference's loader)."""
    import os
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    for m in (rendering.coarse_model, rendering.fine_model):      # device scratch is not part of the state
        if hasattr(m, '_pack'):
            m._pack, m._pack_key = None, None
    torch.save({'rendering': rendering, 'data_config': data_config,
                'Rs_per_ds': rendering.Rs_per_ds if Rs_per_ds is None else Rs_per_ds,
                'seconds_per_dt': seconds_per_dt, 'ref_time': ref_time}, path)
