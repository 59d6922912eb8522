"""Field networks with the reference's constructor signatures, forward contract and state_dict keys
(sunerf/model/model.py, sunerf/model/stellar_model.py), evaluated by the sm_100a kernels.

state_dict keys kept: in_layer.0.freq_bands, in_layer.1.{weight,bias}, layers.{0..6}.{weight,bias},
out_layer.{weight,bias}, log_absortpion.{94,...,335}, volumetric_constant.  Modules are constructed in the
reference's order, so `torch.manual_seed(s)` gives bit-identical initial weights.

Extra (non-reference) knob: `precision` =
  'fp32'  FFMA SIMT kernels: the 1e-5 parity mode for ANY network shape;
  'x3'    (alias 'fp32_tc') the same 1e-5 / 1e-3 gates on tcgen05 tensor cores for the default 8 x 512 network: every operand
          a (hi, lo) fp16 pair, three MMAs per product (csrc/snf_mlp_x3.cu);
  'bf16'  (aliases 'f16', 'fp16') the task's "bf16-MLP mode" on tcgen05: 1e-2 on intensities, 1e-3 on gradients; its MMA
          operands are fp16 since round 2.
Default from $SUNERF_B200_PRECISION, else 'fp32'.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
from torch import nn

from . import ops
from ._lib import SnfError


def canonical_precision(p: str) -> str:
    p = str(p).lower()
    if p in ('f16', 'fp16', 'tc'):
        p = 'bf16'
    if p in ('fp32_tc', 'fp32x3'):
        p = 'x3'
    if p not in ops.MLP_MODES:
        raise ValueError(f'precision must be one of fp32, x3 (fp32_tc), bf16 - got {p}')
    return p


def default_precision() -> str:
    return canonical_precision(os.environ.get('SUNERF_B200_PRECISION', 'fp32'))


class Sine(nn.Module):
    """model.py:66-72.  The field-network kernels fuse it; called on its own it is the reference's one torch op."""

    def __init__(self, w0: float = 1.):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


class PositionalEncoding(nn.Module):
    """model.py:92-132: holds `freq_bands` (buffer, state_dict key).  The field-network kernels fuse the encoding; called
    on its own (model.py:123-132) it is the reference's torch expression: [x, sin(x f / s) (f major), cos(same)]."""

    def __init__(self, d_input: int, n_freqs: int, scale_factor: float = 2., log_space: bool = True):
        super().__init__()
        if not log_space:
            raise SnfError('only log-space frequency bands are built (the reference default)')
        self.d_input, self.n_freqs, self.log_space = d_input, n_freqs, log_space
        self.d_output = d_input * (1 + 2 * n_freqs)
        self.register_buffer('freq_bands', 2. ** torch.linspace(0., n_freqs - 1, n_freqs))
        self.scale_factor = scale_factor

    def forward(self, x):
        arg = x[:, None, :] * self.freq_bands[None, :, None] / self.scale_factor
        return torch.cat([x, torch.sin(arg).reshape(x.shape[0], -1), torch.cos(arg).reshape(x.shape[0], -1)], dim=-1)


class _FieldMLP(torch.autograd.Function):
    """x[M,4] -> raw[M,2]; saves the layer activations in a workspace for the analytic backward."""

    @staticmethod
    def forward(ctx, x, owner, train, off0, off1, *params):
        weights, biases = list(params[0::2]), list(params[1::2])
        mode = owner.precision
        packed = owner._packed_ptr(weights, biases) if mode in ops.TC_MODES else None
        out, ws = ops.mlp_forward(x, weights, biases, (off0, off1), mode=mode, train=train, packed_ptr=packed)
        if train:
            ctx.ws, ctx.packed, ctx.owner = ws, packed, owner
            ctx.save_for_backward(x, *weights)
        return out

    @staticmethod
    def backward(ctx, g):
        x, *weights = ctx.saved_tensors
        gws = [torch.empty_like(w) for w in weights]
        gbs = [torch.empty(w.shape[0], device=w.device, dtype=torch.float32) for w in weights]
        ops.mlp_backward(x, weights, g.contiguous(), ctx.ws, gws, gbs, packed_ptr=ctx.packed)
        ctx.ws = None
        flat = []
        for gw, gb in zip(gws, gbs):
            flat += [gw, gb]
        return (None, None, None, None, None, *flat)


class NeRF(nn.Module):
    """model.py:7-57.  forward(x[M,4]) -> {'inferences': raw[M,2]}."""

    def __init__(self, d_input: int = 4, d_output: int = 2, n_layers: int = 8, d_filter: int = 512,
                 skip: Tuple[int] = (), encoding='positional', precision: str = None):
        super().__init__()
        if d_input != 4 or d_output != 2 or encoding != 'positional' or len(tuple(skip)) != 0:
            raise SnfError('kernels are built for d_input=4, d_output=2, positional encoding, no skips '
                           '(the configuration every shipped reference config uses)')
        self.d_input, self.skip = d_input, skip
        self.act = Sine()
        enc = PositionalEncoding(d_input=d_input, n_freqs=10)
        self.in_layer = nn.Sequential(enc, nn.Linear(enc.d_output, d_filter))
        self.layers = nn.ModuleList([nn.Linear(d_filter, d_filter) for _ in range(n_layers - 1)])
        self.out_layer = nn.Linear(d_filter, d_output)
        self.precision = default_precision() if precision is None else canonical_precision(precision)
        self._pack = None          # (raw tensor, aligned ptr)
        self._pack_key = None

    # -- helpers
    def linear_params(self):
        lins = [self.in_layer[1]] + list(self.layers) + [self.out_layer]
        out = []
        for lin in lins:
            out += [lin.weight, lin.bias]
        return out

    def _out_offsets(self):
        return 0.0, 0.0

    def _packed_ptr(self, weights, biases):
        key = tuple((t.data_ptr(), t._version) for t in list(weights) + list(biases))
        if self._pack is None or self._pack[0].device != weights[0].device:
            self._pack = ops.alloc_packed(weights[0].device)
            self._pack_key = None
        if key != self._pack_key:
            ops.mlp_pack_bf16(weights, biases, self._pack[1])
            self._pack_key = key
        return self._pack[1]

    def raw(self, x: torch.Tensor) -> torch.Tensor:
        off0, off1 = self._out_offsets()
        params = self.linear_params()
        # activations are only kept (and the workspace only sized for them) when a backward can follow
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _FieldMLP.apply(x.reshape(-1, 4), self, train, off0, off1, *params)

    def forward(self, x: torch.Tensor):
        return {'inferences': self.raw(x)}


class EmissionModel(NeRF):
    """model.py:60-63"""

    def __init__(self, **kwargs):
        super().__init__(d_input=4, d_output=2, **kwargs)


class NeRF_DT(NeRF):
    """model.py:136-187: + base offsets, 7 per-wavelength absorption scalars, volumetric constant."""

    def __init__(self, d_input: int = 4, d_output: int = 2, n_layers: int = 8, d_filter: int = 512,
                 skip: Tuple[int] = (), encoding='positional', base_log_temperature: float = 5.0,
                 base_log_density: float = 10.0, precision: str = None):
        super().__init__(d_input=d_input, d_output=d_output, n_layers=n_layers, d_filter=d_filter, skip=skip,
                         encoding=encoding, precision=precision)
        self.base_log_temperature = base_log_temperature
        self.base_log_density = base_log_density
        self.log_absortpion = nn.ParameterDict([[str(c), torch.tensor(1.0e-6, dtype=torch.float32)]
                                                for c in ops.AIA_CHANNELS])
        self.volumetric_constant = nn.Parameter(torch.tensor(1.0, dtype=torch.float32, requires_grad=True))

    def _out_offsets(self):
        return float(self.base_log_density), float(self.base_log_temperature)

    def forward(self, x: torch.Tensor):
        return {'inferences': self.raw(x), 'log_abs': self.log_absortpion, 'vol_c': self.volumetric_constant}


SOLRAD_M = 6.957e8


class SimpleStar(nn.Module):
    """stellar_model.py:5-102: analytic hydrostatic star standing in for a trained field (forward only).
    Quantities are plain floats in the units the reference converts to (R_sun, K, cm^-3)."""

    def __init__(self, h0: float = 60.0e6 / SOLRAD_M, T0: float = 1.4e6, R_s: float = 1.02,
                 t_photosphere: float = 5777.0, rho_0: float = 3.0e8):
        super().__init__()
        self.h0, self.T0, self.R_s, self.t_photosphere, self.rho_0 = float(h0), float(T0), float(R_s), float(t_photosphere), float(rho_0)
        self.log_absortpion = nn.ParameterDict([[str(c), torch.tensor(v, dtype=torch.float32)] for c, v in
                                                zip(ops.AIA_CHANNELS, (20.4, 20.2, 20.0, 19.8, 19.6, 19.4, 19.2))])
        self.stellar_parameters = nn.ParameterDict([['Rs', torch.tensor(self.R_s, dtype=torch.float32)],
                                                    ['h0', torch.tensor(self.h0, dtype=torch.float32)],
                                                    ['T0', torch.tensor(self.T0, dtype=torch.float32)],
                                                    ['rho_0', torch.tensor(self.rho_0, dtype=torch.float32)]])
        self.volumetric_constant = nn.Parameter(torch.tensor(1.0, dtype=torch.float32, requires_grad=True))

        self._consts, self._consts_key = None, None

    def _star_consts(self):
        # the reference computes with the float32 parameter tensors; read them back once per change (one sync)
        sp = self.stellar_parameters
        key = tuple((sp[k].data_ptr(), sp[k]._version) for k in ('rho_0', 'h0', 'T0', 'Rs'))
        if key != self._consts_key:
            self._consts = tuple(float(sp[k].detach().cpu()) for k in ('rho_0', 'h0', 'T0', 'Rs'))
            self._consts_key = key
        return self._consts

    def forward(self, query_points):
        rho_0, h0, T0, Rs = self._star_consts()
        raw = ops.simple_star(query_points.reshape(-1, 4), rho_0, h0, T0, Rs, self.t_photosphere)
        return {'inferences': raw, 'log_abs': self.log_absortpion, 'vol_c': self.volumetric_constant}
