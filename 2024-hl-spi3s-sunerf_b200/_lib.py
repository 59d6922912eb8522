"""ctypes binding of libsunerf_b200.so (the C ABI declared in include/sunerf_b200.h).

The shared library is built in-tree by `build()` (plain nvcc, sm_100a only) and loaded lazily.  There is no
CPU fallback and no alternative backend: if the library is missing or a call fails, the caller gets an
exception.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, 'csrc')
LIB_DIR = os.path.join(_PKG, 'lib')
# SNF_LIB_NAME: developer-only, lets experiment builds (SNF_NVCC_EXTRA) live next to the product library
LIB_PATH = os.path.join(LIB_DIR, os.environ.get('SNF_LIB_NAME', 'libsunerf_b200.so'))
SOURCES = ['snf_sampling.cu', 'snf_rays.cu', 'snf_composite.cu', 'snf_mlp_f32.cu', 'snf_mlp_bf16.cu', 'snf_mlp_bf16_bwd.cu',
           'snf_optim.cu', 'snf_render.cu', 'snf_mlp_x3.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']

_lock = threading.Lock()
_lib = None


class SnfError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise SnfError('nvcc not found: cannot build libsunerf_b200.so')


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(os.path.dirname(_PKG), 'include', 'sunerf_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into lib/libsunerf_b200.so (cross-compiles without a GPU)."""
    with _lock:
        if not force and not _stale():
            return LIB_PATH
        os.makedirs(LIB_DIR, exist_ok=True)
        srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
        tmp = LIB_PATH + '.tmp.%d' % os.getpid()
        # SNF_NVCC_EXTRA: developer-only extra flags (e.g. -DSNF_PROF for the in-kernel cycle counters)
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get('SNF_NVCC_EXTRA', '').split() + ['-o', tmp] + srcs
        if verbose:
            print(' '.join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise SnfError('nvcc failed:\n' + r.stdout + r.stderr)
        os.replace(tmp, LIB_PATH)
        return LIB_PATH


_c = ctypes
_P, _I, _L, _F, _D = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_double
_PROTOS = {
    'snf_version': (_I, []),
    'snf_error_string': (_c.c_char_p, [_I]),
    'snf_launch_count': (_L, []),
    'snf_count_launches': (None, [_L]),
    'snf_config_reserve_sms': (_I, [_I]),
    'snf_stratified_sample': (_I, [_P, _P, _P, _P, _L, _I, _F, _F, _P, _P, _P]),
    'snf_spherical_sample': (_I, [_P, _P, _P, _P, _L, _I, _F, _F, _P, _P, _P]),
    'snf_hier_resample': (_I, [_P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P]),
    'snf_hier_resample_perturb': (_I, [_P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P]),
    'snf_image_rays': (_I, [_P, _I, _I, _D, _D, _D, _D, _L, _L, _P, _P, _P]),
    'snf_make_query': (_I, [_P, _P, _P, _P, _L, _I, _P, _P]),
    'snf_mlp_ws_bytes': (_L, [_L, _I, _I, _I, _I]),
    'snf_mlp_fwd_f32': (_I, [_P, _L, _P, _P, _I, _I, _F, _F, _P, _P, _I, _P]),
    'snf_mlp_bwd_f32': (_I, [_P, _L, _P, _I, _I, _P, _P, _P, _P, _P]),
    'snf_mlp_pack_bytes': (_L, []),
    'snf_mlp_pack_bf16': (_I, [_P, _P, _P, _P]),
    'snf_mlp_fwd_bf16': (_I, [_P, _L, _P, _F, _F, _P, _P, _I, _P]),
    'snf_mlp_bwd_bf16': (_I, [_P, _L, _P, _P, _P, _P, _P, _P]),
    'snf_mlp_fwd_x3': (_I, [_P, _L, _P, _F, _F, _P, _P, _I, _P]),
    'snf_mlp_bwd_x3': (_I, [_P, _L, _P, _P, _P, _P, _P, _P]),
    'snf_simple_star_fwd': (_I, [_P, _L, _F, _F, _F, _F, _F, _P, _P]),
    'snf_composite_emission_fwd': (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P]),
    'snf_composite_emission_bwd': (_I, [_P, _P, _P, _L, _I, _P, _P, _P, _P]),
    'snf_composite_dt_fwd': (_I, [_P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P]),
    'snf_composite_dt_bwd': (_I, [_P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P]),
    'snf_render_epilogue': (_I, [_P, _P, _P, _P, _P, _L, _I, _F, _I, _P, _P, _P, _F, _P, _P]),
    'snf_train_loss': (_I, [_P, _P, _P, _P, _L, _I, _L, _I, _F, _F, _F, _P, _P, _P, _P, _P]),
    'snf_adam_step_sched': (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _P, _P, _P]),
    'snf_debug_time_backward': (_I, [_I]),
    'snf_debug_backward_ms': (_I, [_P]),
    'snf_render_ws_bytes': (_L, [_P, _L, _I]),
    'snf_render_fused_fwd': (_I, [_P, _P, _P, _P, _P, _P, _L, _P, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    'snf_render_fused_bwd': (_I, [_P, _P, _P, _L, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'snf_adam_step': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _L, _F, _F, _P, _P, _P]),
}
EXPORTS = tuple(_PROTOS)


class RenderDesc(ctypes.Structure):
    """struct snf_render_desc of include/sunerf_b200.h"""
    _fields_ = [('kind', _I), ('mode', _I), ('n_hidden', _I), ('d_filter', _I),
                ('W_coarse', _P), ('B_coarse', _P), ('W_fine', _P), ('B_fine', _P),
                ('packed_coarse', _P), ('packed_fine', _P),
                ('out_offset0', _F), ('out_offset1', _F),
                ('t_vals', _P), ('u', _P), ('S', _I), ('n_new', _I),
                ('distance', _F), ('solar_R', _F), ('reg_radius', _F),
                ('C', _I), ('pixel_intensity_factor', _F),
                ('log_abs_coarse', _P), ('vol_c_coarse', _P), ('log_abs_fine', _P), ('vol_c_fine', _P),
                ('table_x', _P), ('table_y', _P)]


def lib() -> ctypes.CDLL:
    """Load (building first if the in-tree .so is missing or older than its sources)."""
    global _lib
    if _lib is None:
        if _stale():
            build()
        with _lock:
            if _lib is None:
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in _PROTOS.items():
                    fn = getattr(L, name)      # AttributeError if the ABI lost a symbol: fail loudly
                    fn.restype, fn.argtypes = res, args
                if L.snf_version() != 100:
                    raise SnfError('libsunerf_b200.so ABI version mismatch')
                _lib = L
    return _lib


def check(code: int, what: str = '') -> None:
    if code != 0:
        msg = lib().snf_error_string(code)
        raise SnfError(f'{what}: {msg.decode() if msg else code} (code {code})')
