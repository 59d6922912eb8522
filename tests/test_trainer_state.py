"""RayTrainer optimiser / schedule state: the reference resumes with trainer.fit(ckpt_path='last')
(sunerf/run_emission.py:38,75), where Lightning restores the Adam moments, step counts and the ExponentialLR state
(sunerf/model/sunerf.py:30-40) next to the weights.  A resumed RayTrainer must continue exactly where it stopped, and its
optimiser state must be loadable by torch.optim.Adam itself (the layout Lightning stores)."""
import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401  (puts the repo root on sys.path, also in spawned workers)

pytestmark = pytest.mark.gpu


def _setup(seed=0, graph=False, lr=1e-3):
    import sunerf_b200 as s
    torch.manual_seed(seed)
    r = s.EmissionRadiativeTransfer(Rs_per_ds=1, model_config={'d_filter': 64, 'n_layers': 3}).cuda()   # fp32 mode: deterministic
    return r, s.RayTrainer(r, lr=lr, use_cuda_graph=graph)


def _batches(n_steps, n=96):
    import sunerf_b200 as s
    out = []
    for i in range(n_steps):
        b = s.rays.synthetic_rays(n, seed=50 + i, H=48, W=48, plate_arcsec=50.0)
        b['t_rand'] = torch.rand(n, 64, generator=torch.Generator().manual_seed(70 + i))
        out.append({k: v.cuda() for k, v in b.items()})
    return out


def _run(tr, batches):
    return [tr.step(b['rays_o'], b['rays_d'], b['times'], b['target'], t_rand=b['t_rand'])['losses'][0].item() for b in batches]


@pytest.mark.parametrize('graph', [False, True])
def test_resume_continues_bit_for_bit(tmp_path, graph):
    bs = _batches(8)
    _, ta = _setup(graph=graph)
    la = _run(ta, bs)                                                  # 8 uninterrupted steps
    rb, tb = _setup(graph=graph)
    lb = _run(tb, bs[:4])
    path = str(tmp_path / 'last.ckpt')
    tb.save_checkpoint(path)
    rc, tc = _setup(seed=123, graph=graph)                             # different initial weights: everything comes from the file
    ck = tc.load_checkpoint(path)
    assert ck['global_step'] == 4 and tc.step_count == 4 and abs(tc.lr - tb.lr) < 1e-18
    lc = _run(tc, bs[4:])
    assert lb + lc == la                                               # fp32 mode is deterministic: same losses to the bit
    assert torch.equal(tc.flat, ta.flat) and torch.equal(tc.exp_avg, ta.exp_avg) and torch.equal(tc.exp_avg_sq, ta.exp_avg_sq)
    sd = tc.state_dict()                                               # the flat form round-trips too
    _, td = _setup(seed=5, graph=graph)
    td.r.load_state_dict(tc.r.state_dict())
    td.load_state_dict(sd)
    assert td.step_count == 8 and torch.equal(td.exp_avg, tc.exp_avg)


def test_optimizer_state_is_torch_adam_layout():
    """checkpoint['optimizer_states'][0] must be what torch.optim.Adam(rendering.parameters()).state_dict() holds: load it
    into a real Adam + ExponentialLR, apply the next step there, and compare with the fused kernel's next step."""
    bs = _batches(4)
    r, tr = _setup()
    _run(tr, bs[:3])
    ref_params = [torch.nn.Parameter(p.detach().cpu().clone()) for p in r.parameters() if p.requires_grad]
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=tr.gamma)
    opt.load_state_dict(tr.optimizer_state_dict())
    sched.load_state_dict(tr.lr_scheduler_state_dict())
    assert abs(opt.param_groups[0]['lr'] - tr.lr) < 1e-18 and sched.last_epoch == 3
    assert all(int(opt.state[p]['step']) == 3 for p in ref_params)
    _run(tr, bs[3:])                                                   # 4th step on the GPU; its gradients feed the torch Adam
    for p_ref, p in zip(ref_params, [p for p in r.parameters() if p.requires_grad]):
        p_ref.grad = tr.grad_view[id(p)].detach().cpu().clone()
    torch.nn.utils.clip_grad_norm_(ref_params, 0.5)
    opt.step(); sched.step()
    for p_ref, p in zip(ref_params, [p for p in r.parameters() if p.requires_grad]):
        assert torch.allclose(p.detach().cpu(), p_ref.detach(), rtol=0, atol=2e-7), (p.detach().cpu() - p_ref).abs().max()
    assert abs(opt.param_groups[0]['lr'] - tr.lr) < 1e-15
    # and back: a state saved by torch's Adam loads into a fresh trainer
    r2, t2 = _setup(seed=9)
    r2.load_state_dict(r.state_dict())
    t2.load_optimizer_state_dict(opt.state_dict())
    assert t2.step_count == 4 and abs(t2.lr - opt.param_groups[0]['lr']) < 1e-18
    for p_ref, p in zip(ref_params, [p for p in r2.parameters() if p.requires_grad]):
        o, n = t2.segment[id(p)]
        assert torch.equal(t2.exp_avg[o:o + n].cpu().view(p.shape), opt.state[p_ref]['exp_avg'])


def test_moved_module_is_detected():
    import sunerf_b200 as s
    r, tr = _setup()
    b = _batches(1)[0]
    _run(tr, [b])
    r.double()                                                         # re-types (re-allocates) every parameter
    with pytest.raises(s.SnfError):
        _run(tr, [b])
