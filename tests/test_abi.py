"""CPU suite: the C-ABI library builds for sm_100a without a GPU, loads, and exports exactly what include/*.h declares."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'sunerf_b200.h')).read()
    src += open(os.path.join(ROOT, 'include', 'sunerf_b200_debug.h')).read()      # measurement aids, same library
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(snf_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_header_symbols():
    import sunerf_b200
    path = sunerf_b200.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/sunerf_b200.h but not exported'
    assert set(names) == set(sunerf_b200._lib.EXPORTS), set(names) ^ set(sunerf_b200._lib.EXPORTS)
    assert sunerf_b200._lib.lib().snf_version() == 100
    assert sunerf_b200._lib.lib().snf_error_string(-2).decode().startswith('shape')


def test_no_cpu_fallback():
    import sunerf_b200
    r = sunerf_b200.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'n_layers': 2, 'd_filter': 16})
    with pytest.raises(sunerf_b200.SnfError):
        r(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 1))
    with pytest.raises(sunerf_b200.SnfError):
        sunerf_b200.ops.stratified_sample(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(64), None, 1.3, 1.0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, '2024-hl-spi3s-sunerf_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            assert 'oracle' not in open(os.path.join(pkg, f)).read().replace('# oracle', ''), f


def test_reference_surface():
    import inspect
    import sunerf_b200 as s
    sig = inspect.signature(s.SuNeRFRendering.__init__)
    assert list(sig.parameters)[1:] == ['Rs_per_ds', 'sampling_config', 'hierarchical_sampling_config', 'model', 'model_config']
    sig = inspect.signature(s.DensityTemperatureRadiativeTransfer.__init__)
    assert {'model_config', 'device', 'aia_exp_time', 'pixel_intensity_factor'} <= set(sig.parameters)
    with pytest.raises(ValueError):
        s.EmissionRadiativeTransfer(Rs_per_ds=1, sampling_config={'type': 'nope'})
    with pytest.raises(ValueError):
        s.EmissionRadiativeTransfer(Rs_per_ds=1, hierarchical_sampling_config={'type': 'nope'})
    m = s.NeRF_DT()
    keys = set(m.state_dict().keys())
    assert {'in_layer.0.freq_bands', 'in_layer.1.weight', 'layers.6.bias', 'out_layer.weight', 'log_absortpion.94',
            'log_absortpion.335', 'volumetric_constant'} <= keys
    assert sum(p.numel() for p in s.NeRF().parameters()) == 1883138


def test_render_workspace_size_is_host_arithmetic():
    """snf_render_ws_bytes / snf_mlp_ws_bytes need no GPU: sizes grow with the batch, a training forward keeps more than an
    inference one, and a bad descriptor is an error code, not a crash."""
    import sunerf_b200
    from sunerf_b200 import _lib
    L = _lib.lib()
    d = _lib.RenderDesc()
    d.kind, d.mode, d.n_hidden, d.d_filter = 0, 1, 8, 512
    d.packed_coarse = d.packed_fine = 1024          # any non-null value: nothing is dereferenced
    d.t_vals = d.u = 1024
    d.S, d.n_new = 64, 128
    small = L.snf_render_ws_bytes(ctypes.byref(d), 256, 0)
    big = L.snf_render_ws_bytes(ctypes.byref(d), 1024, 0)
    train = L.snf_render_ws_bytes(ctypes.byref(d), 1024, 1)
    assert 0 < small < big < train
    assert train >= L.snf_mlp_ws_bytes(1024 * 64, 8, 512, 1, 1) + L.snf_mlp_ws_bytes(1024 * 192, 8, 512, 1, 1)
    d.S, d.n_new = 200, 128                          # S + n_new > 256: outside what the compositing kernels are built for
    assert L.snf_render_ws_bytes(ctypes.byref(d), 256, 0) == -2
    d.S, d.kind = 64, 7
    assert L.snf_render_ws_bytes(ctypes.byref(d), 256, 0) == -1
    assert L.snf_render_ws_bytes(None, 256, 0) == -1


def test_headers_are_plain_c():
    """The boundary is a C ABI (`extern "C"`, plain pointers and sizes, no torch / CUDA types in the signatures): both headers
    compile as C99 with nothing but the standard headers, and a C translation unit can take the address of every entry."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc')
    inc = os.path.join(ROOT, 'include')
    for h in ('sunerf_b200.h', 'sunerf_b200_debug.h'):
        r = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', os.path.join(inc, h)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    body = '#include "sunerf_b200.h"\n#include "sunerf_b200_debug.h"\nconst void *table[] = {\n' + \
           ''.join(f'  (const void *)&{n},\n' for n in _declared()) + '};\nint main(void) { return table[0] == 0; }\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, 'use.c')
        open(src, 'w').write(body)
        r = subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', inc, '-c', src, '-o', os.path.join(d, 'use.o')],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
