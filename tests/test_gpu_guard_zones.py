"""Out-of-bounds WRITE check of every kernel on the path (-m gpu).

compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_closed_on_pool.txt), so this is the memcheck the
repo can run itself: every device buffer the host layer hands to the C ABI (`torch.empty` / `torch.empty_like` inside
sunerf_b200.ops and sunerf_b200.fused: kernel outputs, the field-network workspaces with the saved activation images)
is carved out of a larger allocation whose 8 KB on either side are filled with a sentinel byte; after a train step /
render at RAGGED sizes (partial tiles, partial warps, partial 128-point images) in every mode, every guard zone must
still hold the sentinel.  The results are checked against the un-guarded run too (same seeds -> same bits in fp32
mode), so an in-bounds but misplaced write shows as well."""
import pytest
import torch

from conftest import ROOT  # noqa: F401

pytestmark = pytest.mark.gpu

GUARD = 8192
SENTINEL = 0xA5


class _GuardedTorch:
    """Stands in for the `torch` module inside ops.py / fused.py: `empty` and `empty_like` on a CUDA device come with guard zones."""

    def __init__(self, real):
        self._real = real
        self.allocs = []          # (raw uint8 buffer, payload bytes, label)

    def __getattr__(self, name):
        return getattr(self._real, name)

    def _guarded(self, shape, device, dtype):
        real = self._real
        dtype = dtype or real.float32
        shape = tuple(int(v) for v in shape)
        n = 1
        for v in shape:
            n *= v
        nbytes = n * real.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 16
        raw = real.full((2 * GUARD + nbytes + pad,), SENTINEL, device=device, dtype=real.uint8)
        self.allocs.append((raw, nbytes, f'{shape} {dtype}'))
        return raw[GUARD:GUARD + nbytes].view(dtype).view(shape)

    def empty(self, *shape, device=None, dtype=None, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, self._real.Size)):
            shape = tuple(shape[0])
        if device is None or self._real.device(device).type != 'cuda':
            return self._real.empty(*shape, device=device, dtype=dtype, **kw)
        return self._guarded(shape, device, dtype)

    def empty_like(self, t, **kw):
        if not t.is_cuda or kw:
            return self._real.empty_like(t, **kw)
        return self._guarded(t.shape, t.device, t.dtype)

    def check(self):
        self._real.cuda.synchronize()
        bad = []
        for raw, nbytes, label in self.allocs:
            lo, hi = raw[:GUARD], raw[GUARD + nbytes:]
            if not (bool((lo == SENTINEL).all()) and bool((hi == SENTINEL).all())):
                bad.append(label)
        return bad


@pytest.fixture
def guarded(monkeypatch):
    import sunerf_b200.ops as ops
    import sunerf_b200.fused as fused
    g = _GuardedTorch(torch)
    monkeypatch.setattr(ops, 'torch', g)
    monkeypatch.setattr(fused, 'torch', g)
    return g


def _emission(precision, seed=5):
    import sunerf_b200 as s
    torch.manual_seed(seed)
    return s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'precision': precision}).cuda()


def _dt(precision, seed=5):
    import sunerf_b200 as s
    torch.manual_seed(seed)
    return s.DensityTemperatureRadiativeTransfer(Rs_per_ds=1, model=s.NeRF_DT, pixel_intensity_factor=1e17,
                                                 model_config={'precision': precision}).cuda()


def _batch(n, seed):
    import sunerf_b200 as s
    b = {k: v.cuda() for k, v in s.rays.synthetic_rays(n, seed=seed, H=64, W=64, plate_arcsec=38.0).items()}
    b['t_rand'] = torch.rand(n, 64, generator=torch.Generator().manual_seed(seed + 1)).cuda()
    return b


# 1 ray: a single partial tile; 37: partial warp groups, 37 * 192 = 7104 points = 55.5 tiles of 128 (odd tile count ->
# padded tile pair); 301: several rounds of CTA pairs with a ragged tail
@pytest.mark.parametrize('precision', ['fp32', 'x3', 'bf16'])
@pytest.mark.parametrize('n', [1, 37, 301])
def test_emission_train_step_and_render_stay_inside_their_buffers(guarded, precision, n):
    import sunerf_b200 as s
    r = _emission(precision)
    tr = s.RayTrainer(r)
    b = _batch(n, 11)
    for _ in range(2):
        res = tr.step(b['rays_o'], b['rays_d'], b['times'], b['target'], t_rand=b['t_rand'])
    with torch.no_grad():
        out = r(b['rays_o'], b['rays_d'], b['times'], t_rand=b['t_rand'])
    assert len(guarded.allocs) > 20                      # the allocations really went through the guarded allocator
    assert guarded.check() == []
    assert torch.isfinite(res['losses']).all() and torch.isfinite(out['fine_image']).all()
    tr.check_finite()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_density_temperature_train_step_stays_inside_its_buffers(guarded, precision):
    import sunerf_b200 as s
    n = 53
    r = _dt(precision)
    tr = s.RayTrainer(r)
    b = _batch(n, 13)
    wl = torch.tensor([94., 131., 171., 193., 211., 304., 335.]).repeat(n, 1)
    wl[n // 2:] = torch.tensor([0., 0., 171., 193., 211., 304., 0.])
    target = torch.rand(n, 7).cuda()
    for _ in range(2):
        res = tr.step(b['rays_o'], b['rays_d'], b['times'], target, wl.cuda(), t_rand=b['t_rand'])
    assert guarded.check() == []
    assert torch.isfinite(res['losses']).all()
    tr.check_finite()


def test_guarded_run_is_bit_identical_to_the_plain_run_in_fp32(guarded, monkeypatch):
    """Same seeds, with and without guard zones: an in-bounds write to the wrong place would change the result."""
    import sunerf_b200 as s
    import sunerf_b200.ops as ops
    import sunerf_b200.fused as fused
    b = _batch(45, 17)
    r = _emission('fp32')
    with torch.no_grad():
        a = r(b['rays_o'], b['rays_d'], b['times'], t_rand=b['t_rand'])
    assert guarded.check() == []
    monkeypatch.setattr(ops, 'torch', torch)
    monkeypatch.setattr(fused, 'torch', torch)
    with torch.no_grad():
        c = r(b['rays_o'], b['rays_d'], b['times'], t_rand=b['t_rand'])
    for k in a:
        assert torch.equal(a[k], c[k]), k


def test_samplers_and_fused_entry_stay_inside_their_buffers(guarded):
    import sunerf_b200 as s
    from sunerf_b200 import ops
    b = _batch(77, 19)
    r = _emission('bf16')
    fr = s.FusedRender(r, 77, train=True)
    out = fr.forward(b['rays_o'], b['rays_d'], b['times'], t_rand=b['t_rand'], reg_grad_scale=1e-3)
    fr.backward(torch.ones_like(out['coarse_image']), torch.ones_like(out['fine_image']))
    # per-ray uniform draws (HierarchicalSampler(perturb=True)) and the spherical sampler
    hs = s.HierarchicalSampler(128, perturb=True).cuda()
    z, _ = r.sampler.sample_z(b['rays_o'], b['rays_d'], t_rand=b['t_rand'])
    w = torch.rand_like(z)
    hs.resample(z, w)
    sp = s.SphericalSampler(Rs_per_ds=1, n_samples=64, perturb=True).cuda()
    sp(b['rays_o'], b['rays_d'])
    assert guarded.check() == []
