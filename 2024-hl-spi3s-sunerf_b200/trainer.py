"""Training fast path: one ray batch -> loss -> analytic backward -> (NCCL all-reduce) -> clip + Adam, without
autograd and without a host sync.  Same semantics as the reference's Lightning step
(sunerf/model/sunerf.py:98-131 emission, :173-206 density-temperature; optimiser :30-40; clip run_emission.py:72),
with Lightning's single-process 'dp' strategy (run_emission.py:69) replaced by one process per GPU, rays
sharded by rank and ONE all-reduce of the flat fp32 gradient buffer per step.

All parameters of the rendering module are re-homed as views into one flat buffer (16-byte aligned
segments), so the kernels write gradients in place, NCCL reduces one tensor and the optimiser is one
pass.  `rendering.state_dict()` keeps the reference's keys and shapes.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, Optional

import torch

from . import _lib, ops, parallel
from ._lib import SnfError
from .rendering import DensityTemperatureRadiativeTransfer, SuNeRFRendering


class ImageAsinhScaling(torch.nn.Module):
    """sunerf/train/scaling.py:17-28 (plain torch ops; the fast path fuses it into the loss kernel)."""

    def __init__(self, vmax=1, a=0.005):
        super().__init__()
        import numpy as np
        self.normalization = torch.nn.Parameter(torch.tensor(np.arcsinh(1 / a), dtype=torch.float32), requires_grad=False)
        self.a = torch.nn.Parameter(torch.tensor(a, dtype=torch.float32), requires_grad=False)
        self.vmax = torch.nn.Parameter(torch.tensor(vmax, dtype=torch.float32), requires_grad=False)

    def forward(self, image):
        return torch.asinh(image / self.vmax / self.a) / self.normalization


def _align4(n: int) -> int:
    return (n + 3) // 4 * 4


class RayTrainer:
    def __init__(self, rendering: SuNeRFRendering, lr: float = 1e-4, lr_end: float = 1e-5, lr_iterations: float = 1e6,
                 clip_norm: float = 0.5, lambda_image: float = 1.0, lambda_regularization: float = 1.0,
                 asinh_a: float = 0.005, process_group=None, device=None, use_cuda_graph: bool = False):
        self.r = rendering
        self.dt = isinstance(rendering, DensityTemperatureRadiativeTransfer)
        self.dev = torch.device(device) if device is not None else next(rendering.parameters()).device
        if self.dev.type != 'cuda':
            raise SnfError('RayTrainer needs the rendering module on a CUDA device')
        self.lr, self.gamma, self.clip = lr, (lr_end / lr) ** (1 / lr_iterations), clip_norm
        self.lam_img, self.lam_reg, self.asinh_a = lambda_image, lambda_regularization, asinh_a
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (
            torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        # multi-GPU schedule switches.  Measured on 8 GPUs of one box against the 1-GPU step of the same
        # box (3.40 ms): both buckets exchanged after the fine backward, no SM reserve - 3.57 ms (the default); coarse bucket
        # exchanged early from the side stream on 4 SMs left free by the persistent grids - 3.79 ms: NCCL's CTAs that do not
        # fit the reserve wait for a tail of the fine pass while their peers spin.  On 2 GPUs the variants are equal (3.53 ms).
        import os
        self.early_coarse_reduce = os.environ.get('SNF_EARLY_REDUCE', '0') == '1'
        self.reserve_sms = int(os.environ.get('SNF_RESERVE_SMS', '0')) if self.world > 1 else 0
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().snf_config_reserve_sms(self.reserve_sms), 'snf_config_reserve_sms')
        self.step_count = 0
        self.lr0 = lr
        self._flatten()
        if self.world > 1:
            # replicas must start identical: rank 0's parameters win (a differently seeded or differently restored rank
            # would otherwise diverge silently while the all-reduce keeps succeeding)
            torch.distributed.broadcast(self.flat, src=torch.distributed.get_global_rank(process_group, 0)
                                        if process_group is not None else 0, group=process_group)
        self.scratch = torch.zeros(1024, device=self.dev)
        self.grad_norm = torch.zeros(1, device=self.dev)
        self.finite_flag = torch.zeros(1, device=self.dev, dtype=torch.int32)
        # CUDA-graph replay of the whole step (the NCCL all-reduces of the multi-GPU step are captured with it).
        # lr / step live on the device so that nothing on the host changes between replays.
        self.use_cuda_graph = bool(use_cuda_graph)
        self.overlap_backward = True                   # coarse backward on a side stream beside the fine pass (off: one kernel at a time)
        self._side_stream = torch.cuda.Stream(device=self.dev)
        self._zero1 = torch.zeros(1, device=self.dev)
        self._device_sched = self.use_cuda_graph       # once the schedule lives on the device it stays there (eager steps too)
        self.sched = torch.tensor([lr, 1.0, self.gamma, 5e-5], device=self.dev, dtype=torch.float64)
        self._graph, self._g_in, self._g_out, self._g_key, self._g_warm = None, None, None, None, 0

    # -- flat parameter / gradient buffers: [fine model | coarse model], log_abs contiguous in channel order
    def _flatten(self):
        segs = []   # (param, offset)
        off = 0
        self.model_range = {}
        for name in ('fine_model', 'coarse_model'):
            model = getattr(self.r, name)
            start = off
            for p in model.linear_params():
                segs.append((p, off)); off += _align4(p.numel())
            if self.dt:
                self.__dict__[f'la_off_{name}'] = off
                for c in ops.AIA_CHANNELS:
                    segs.append((model.log_absortpion[str(c)], off)); off += 1
                off = _align4(off)
                self.__dict__[f'vc_off_{name}'] = off
                segs.append((model.volumetric_constant, off)); off += 4
            self.model_range[name] = (start, off)
        known = {id(p) for p, _ in segs}
        extra = [p for p in self.r.parameters() if id(p) not in known and p.requires_grad]
        if extra:
            raise SnfError('rendering module has trainable parameters the fast path does not know')
        self.n_flat = off
        self.flat = torch.zeros(off, device=self.dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(off, device=self.dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(off, device=self.dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=self.dev, dtype=torch.float32)
        self.grad_view = {}
        self.segment = {}          # id(parameter) -> (offset, numel) in the flat buffers
        self._segs = segs
        with torch.no_grad():
            for p, o in segs:
                n = p.numel()
                self.flat[o:o + n].copy_(p.detach().reshape(-1).to(self.dev))
                p.data = self.flat[o:o + n].view(p.shape)
                self.grad_view[id(p)] = self.flat_grad[o:o + n].view(p.shape)
                self.segment[id(p)] = (o, n)

    def _check_homed(self):
        """The parameters are views into `self.flat`; `rendering.to(device)`, `.half()`, `.float()` or a checkpoint load that
        REPLACES `p.data` would silently detach them (the kernels would keep training the flat buffer, the module would
        keep its stale copy).  Checked once per step on the first and last segment - a module-wide move changes both."""
        for p, o in (self._segs[0], self._segs[-1]):
            if p.data_ptr() != self.flat.data_ptr() + 4 * o or p.device != self.dev or p.dtype != torch.float32:
                raise SnfError('the rendering module was moved / re-typed / re-allocated after RayTrainer took ownership of its '
                               'parameters (they are views into one flat buffer): build a new RayTrainer, or load weights '
                               'with load_state_dict / copy_ (in place)')

    def _grads(self, model):
        ps = model.linear_params()
        return [self.grad_view[id(p)] for p in ps[0::2]], [self.grad_view[id(p)] for p in ps[1::2]]

    def _field(self, model, query, train=True):
        ps = model.linear_params()
        weights, biases = ps[0::2], ps[1::2]
        packed = model._packed_ptr(weights, biases) if model.precision in ops.TC_MODES else None
        out, ws = ops.mlp_forward(query.view(-1, 4), weights, biases, model._out_offsets(), mode=model.precision,
                                  train=train, packed_ptr=packed)
        return out, ws, weights, packed

    @torch.no_grad()
    def step(self, rays_o, rays_d, times, target, wavelengths=None, t_rand=None) -> Dict[str, torch.Tensor]:
        """One training step. With use_cuda_graph the step is captured once per batch shape (after two eager warm-up
        steps) and replayed: inputs are copied into static buffers, the returned tensors are static buffers too."""
        # host (pinned) tensors are accepted: eager steps move them to the device, replays copy them straight into the
        # graph's static input buffers
        to_dev = lambda x: x if (x is None or x.is_cuda) else x.to(self.dev, non_blocking=True)
        self._check_homed()
        if torch.cuda.current_device() != self.dev.index:
            with torch.cuda.device(self.dev):          # the C ABI launches on the CURRENT device's stream
                return self.step(rays_o, rays_d, times, target, wavelengths, t_rand)
        if rays_o.shape[0] == 0:
            return self._empty_step()
        if not self.use_cuda_graph:
            return self._step_impl(to_dev(rays_o), to_dev(rays_d), to_dev(times), to_dev(target), to_dev(wavelengths),
                                   to_dev(t_rand), device_sched=self._device_sched)
        if t_rand is None and self.r.sampler.perturb:      # the reference's torch.rand draw, outside the graph
            t_rand = torch.rand((rays_o.shape[0], self.r.sampler.t_vals.shape[1]), device=self.dev)
        ins = {'rays_o': rays_o, 'rays_d': rays_d, 'times': times, 'target': target}
        if wavelengths is not None:
            ins['wavelengths'] = wavelengths
        if t_rand is not None:
            ins['t_rand'] = t_rand
        key = tuple((k, tuple(v.shape)) for k, v in ins.items())
        if key != self._g_key:
            self._graph, self._g_key, self._g_warm = None, key, 0
        if self._graph is None and self._g_warm < 2:       # warm-up: lazy allocations / one-time attribute calls
            self._g_warm += 1
            return self._step_impl(to_dev(rays_o), to_dev(rays_d), to_dev(times), to_dev(target), to_dev(wavelengths),
                                   to_dev(t_rand), device_sched=True)
        if self._graph is None:
            self._g_in = {k: v.detach().to(self.dev, torch.float32).contiguous().clone() for k, v in ins.items()}
            torch.cuda.synchronize(self.dev)
            self._graph = torch.cuda.CUDAGraph()
            l0 = ops.launch_count()
            with torch.cuda.graph(self._graph):
                gi = self._g_in
                self._g_out = self._step_impl(gi['rays_o'], gi['rays_d'], gi['times'], gi['target'], gi.get('wavelengths'),
                                              gi.get('t_rand'), device_sched=True, host_bookkeeping=False)
            self._g_launches = ops.launch_count() - l0     # kernels per replay (they were recorded, not executed)
            _lib.lib().snf_count_launches(-self._g_launches)
            # the capture itself did not execute: fall through to the first replay with the caller's data
        for k, v in ins.items():
            self._g_in[k].copy_(v, non_blocking=True)
        self._graph.replay()
        _lib.lib().snf_count_launches(self._g_launches)
        self._host_bookkeeping()
        return self._g_out

    def _empty_step(self) -> Dict[str, torch.Tensor]:
        """A rank whose shard of a (tail) batch is empty still joins both all-reduces - with zero gradients - and applies the
        same optimiser step as everybody else, so the replicas stay identical and nobody hangs in NCCL."""
        self.flat_grad.zero_()
        # the SAME issue order as _step_impl on the ranks that have rays (collectives pair up by order, and the two buckets
        # have the same size: a swapped order would silently add fine gradients to coarse ones;
        # the late schedule exchanges the whole flat gradient in one call)
        if self.early_coarse_reduce:
            parallel.wait_all([self._reduce_async('coarse_model'), self._reduce_async('fine_model')])
        else:
            parallel.wait_all([parallel.allreduce_sum_async(self.flat_grad, 0, self.n_flat, self.pg)])
        self._optimizer_step(self._device_sched)
        self._host_bookkeeping()
        nan = torch.full((4,), float('nan'), device=self.dev)
        C = 0
        return {'losses': nan, 'psnr': nan[0], 'coarse_image': torch.empty(0, C, device=self.dev),
                'fine_image': torch.empty(0, C, device=self.dev), 'grad_norm': self.grad_norm,
                'z_vals_hierarchical': torch.empty(0, self.r.sampler_hierarchical.n_samples, device=self.dev)}

    def _optimizer_step(self, device_sched: bool):
        # grads averaged over ranks == Lightning dp's mean of replica losses
        if device_sched:
            ops.adam_step_sched(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.sched, self.scratch,
                                self.grad_norm, clip_norm=self.clip, grad_scale=1.0 / self.world)
        else:
            ops.adam_step(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count + 1, self.lr, self.scratch,
                          self.grad_norm, clip_norm=self.clip, grad_scale=1.0 / self.world)

    def _host_bookkeeping(self):
        self.step_count += 1
        for name in ('fine_model', 'coarse_model'):   # packed bf16 weights must be refreshed next forward
            getattr(self.r, name)._pack_key = None
        if self.lr > 5e-5:                             # sunerf.py:36-40
            self.lr *= self.gamma

    def _step_impl(self, rays_o, rays_d, times, target, wavelengths=None, t_rand=None, device_sched=False,
                   host_bookkeeping=True) -> Dict[str, torch.Tensor]:
        r = self.r
        N = rays_o.shape[0]
        z, _ = r.sampler.sample_z(rays_o, rays_d, t_rand=t_rand)
        S = z.shape[1]
        q_c = ops.make_query(rays_o, rays_d, z, times)
        raw_c, ws_c, w_c, pk_c = self._field(r.coarse_model, q_c)
        raw_c = raw_c.view(N, S, 2)
        if self.dt:
            la_c = self.flat[self.la_off_coarse_model:self.la_off_coarse_model + 7]
            vc_c = self.flat[self.vc_off_coarse_model:self.vc_off_coarse_model + 1]
            la_f = self.flat[self.la_off_fine_model:self.la_off_fine_model + 7]
            vc_f = self.flat[self.vc_off_fine_model:self.vc_off_fine_model + 1]
            F = float(r.pixel_intensity_factor)
            img_c, wts_c, _ = ops.composite_dt_fwd(raw_c, z, wavelengths, la_c, vc_c, r._table_x, r._table_y, F)
        else:
            img_c, wts_c, _ = ops.composite_emission_fwd(raw_c, z, rays_d)
        # ---- coarse backward, on a side stream.  The gradient of the coarse image needs the coarse image only, and the
        # resampled depths carry no gradient, so the coarse network's whole backward is independent of the fine pass: its
        # CTAs fill the SMs that the fine pass leaves idle - the last round of each persistent kernel (3.46 / 10.4 rounds
        # of 74 CTA pairs at 1024 rays) and the short HBM-class kernels between the field-network launches.
        main = torch.cuda.current_stream()
        side = self._side_stream if self.overlap_backward else main
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            # coarse term of the loss alone (same kernel, same arithmetic: g_ic is bit-identical to the full call's)
            _, g_ic, _, _ = ops.train_loss(img_c, img_c, target, self._zero1, asinh_scaling=not self.dt,
                                           asinh_a=self.asinh_a, lambda_image=self.lam_img, lambda_reg=0.0,
                                           finite_flag=self.finite_flag)
            gw, gb = self._grads(r.coarse_model)
            if self.dt:
                lo, hi = self.la_off_coarse_model, self.vc_off_coarse_model + 4
                self.flat_grad[lo:hi].zero_()
                g_raw_c, _, _ = ops.composite_dt_bwd(raw_c, z, wavelengths, la_c, vc_c, r._table_x, r._table_y, F, g_ic, None,
                                                     self.flat_grad[lo:lo + 7], self.flat_grad[hi - 4:hi - 3])
            else:
                g_raw_c = ops.composite_emission_bwd(raw_c, z, rays_d, g_ic.view(-1), None)
            ops.mlp_backward(q_c.view(-1, 4), w_c, g_raw_c.view(-1, 2), ws_c, gw, gb, packed_ptr=pk_c)
            # the coarse bucket is exchanged as soon as it exists (a millisecond before the fine backward ends): NCCL's
            # CTAs slot into the tails of the fine pass's persistent kernels instead of queueing up behind the step
            h_coarse = self._reduce_async('coarse_model') if self.early_coarse_reduce else None
        # ---- fine pass
        new_z, z_comb = r.sampler_hierarchical.resample(z, wts_c)
        Sf = z_comb.shape[1]
        q_f = ops.make_query(rays_o, rays_d, z_comb, times)
        raw_f, ws_f, w_f, pk_f = self._field(r.fine_model, q_f)
        raw_f = raw_f.view(N, Sf, 2)
        if self.dt:
            img_f, wts_f, qq = ops.composite_dt_fwd(raw_f, z_comb, wavelengths, la_f, vc_f, r._table_x, r._table_y, F)
        else:
            img_f, wts_f, qq = ops.composite_emission_fwd(raw_f, z_comb, rays_d)
        _, _, reg, g_q = ops.render_epilogue(rays_o, rays_d, z_comb, wts_f, qq, r.reg_radius / r.Rs_per_ds, r.kind,
                                             grad_scale=self.lam_reg / float(N * Sf), want_gq=True)
        losses, _, g_if, _ = ops.train_loss(img_c, img_f, target, reg, asinh_scaling=not self.dt, asinh_a=self.asinh_a,
                                            lambda_image=self.lam_img, lambda_reg=self.lam_reg,
                                            finite_flag=self.finite_flag)
        gw_f, gb_f = self._grads(r.fine_model)
        if self.dt:
            lo, hi = self.la_off_fine_model, self.vc_off_fine_model + 4
            self.flat_grad[lo:hi].zero_()
            g_raw_f, _, _ = ops.composite_dt_bwd(raw_f, z_comb, wavelengths, la_f, vc_f, r._table_x, r._table_y, F, g_if, g_q,
                                                 self.flat_grad[lo:lo + 7], self.flat_grad[hi - 4:hi - 3])
        else:
            g_raw_f = ops.composite_emission_bwd(raw_f, z_comb, rays_d, g_if.view(-1), g_q)
        ops.mlp_backward(q_f.view(-1, 4), w_f, g_raw_f.view(-1, 2), ws_f, gw_f, gb_f, packed_ptr=pk_f)
        if self.early_coarse_reduce:
            h_fine = self._reduce_async('fine_model')
            if side is not main:
                main.wait_stream(side)
        else:
            # late schedule: both buckets are complete here and contiguous - ONE all-reduce of the whole flat gradient
            # (15 MB on NVSwitch is latency-bound: a second NCCL launch costs more than its bytes)
            if side is not main:
                main.wait_stream(side)
            h_fine, h_coarse = parallel.allreduce_sum_async(self.flat_grad, 0, self.n_flat, self.pg), None
        parallel.wait_all([h_coarse, h_fine])
        # ---- optimiser
        self._optimizer_step(device_sched)
        if host_bookkeeping:
            self._host_bookkeeping()
        else:                                          # under capture: the next replay must re-pack the weights
            for name in ('fine_model', 'coarse_model'):
                getattr(r, name)._pack_key = None
        # what the reference logs per step (sunerf.py:121-129): loss, coarse, fine, regularization, psnr = -10 log10(fine)
        psnr = -10. * torch.log10(losses[2])
        return {'losses': losses, 'psnr': psnr, 'coarse_image': img_c, 'fine_image': img_f, 'grad_norm': self.grad_norm,
                'z_vals_hierarchical': new_z}

    def _reduce_async(self, name):
        lo, hi = self.model_range[name]
        return parallel.allreduce_sum_async(self.flat_grad, lo, hi, self.pg)

    def check_finite(self) -> None:
        """The reference asserts on NaN/Inf every step (sunerf.py:105-107); here it is one device flag read on demand."""
        if int(self.finite_flag.item()) != 0:
            raise AssertionError('! [Numerical Alert] a rendered output contains NaN or Inf')

    # ------------------------------------------------------------------ optimiser / schedule state (resume)
    # The reference resumes with `trainer.fit(ckpt_path='last')` (run_emission.py:38,75): Lightning restores the Adam
    # moments and step counts (`optimizer_states`) and the ExponentialLR state (`lr_schedulers`) next to the weights.
    def _sync_sched_from_device(self):
        if self._device_sched:
            lr, step = self.sched[:2].tolist()
            self.lr, self.step_count = float(lr), int(round(step)) - 1

    def state_dict(self) -> Dict[str, Any]:
        """Flat optimiser state of this trainer (moments, step, lr); weights travel in `rendering.state_dict()`."""
        self._sync_sched_from_device()
        return {'exp_avg': self.exp_avg.detach().cpu().clone(), 'exp_avg_sq': self.exp_avg_sq.detach().cpu().clone(),
                'step_count': self.step_count, 'lr': self.lr, 'lr0': self.lr0, 'gamma': self.gamma, 'n_flat': self.n_flat}

    def load_state_dict(self, sd: Dict[str, Any]) -> None:
        if int(sd['n_flat']) != self.n_flat:
            raise SnfError(f"optimiser state is for {sd['n_flat']} flat parameters, this trainer has {self.n_flat}")
        self.exp_avg.copy_(sd['exp_avg']); self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        self._set_schedule(float(sd['lr']), int(sd['step_count']))

    def _set_schedule(self, lr: float, step_count: int):
        self.lr, self.step_count = lr, step_count
        self.sched.copy_(torch.tensor([lr, step_count + 1.0, self.gamma, 5e-5], dtype=torch.float64))
        self._graph, self._g_key = None, None          # a captured step holds no host scalars, but start clean

    def _adam_params(self):
        return [p for p in self.r.parameters() if p.requires_grad]

    def optimizer_state_dict(self) -> Dict[str, Any]:
        """`torch.optim.Adam(rendering.parameters(), lr).state_dict()` layout (sunerf/model/sunerf.py:31), i.e. what
        Lightning stores under checkpoint['optimizer_states'][0]: per-parameter step / exp_avg / exp_avg_sq in
        `rendering.parameters()` order."""
        self._sync_sched_from_device()
        state = {}
        params = self._adam_params()
        if self.step_count > 0:
            for i, p in enumerate(params):
                o, n = self.segment[id(p)]
                state[i] = {'step': torch.tensor(float(self.step_count)),
                            'exp_avg': self.exp_avg[o:o + n].view(p.shape).detach().cpu().clone(),
                            'exp_avg_sq': self.exp_avg_sq[o:o + n].view(p.shape).detach().cpu().clone()}
        group = {'lr': self.lr, 'betas': (0.9, 0.999), 'eps': 1e-8, 'weight_decay': 0, 'amsgrad': False, 'maximize': False,
                 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None, 'decoupled_weight_decay': False,
                 'initial_lr': self.lr0, 'params': list(range(len(params)))}
        return {'state': state, 'param_groups': [group]}

    def load_optimizer_state_dict(self, sd: Dict[str, Any]) -> None:
        params = self._adam_params()
        group = sd['param_groups'][0]
        if len(group['params']) != len(params):
            raise SnfError(f"optimizer state has {len(group['params'])} parameters, the rendering module has {len(params)}")
        if tuple(group.get('betas', (0.9, 0.999))) != (0.9, 0.999) or abs(group.get('eps', 1e-8) - 1e-8) > 0:
            raise SnfError('the fused optimiser implements the reference Adam (betas 0.9/0.999, eps 1e-8)')
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        steps = set()
        for i, p in enumerate(params):
            st = sd['state'].get(i, sd['state'].get(str(i)))
            if st is None:
                continue
            o, n = self.segment[id(p)]
            self.exp_avg[o:o + n].copy_(st['exp_avg'].reshape(-1)); self.exp_avg_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(round(float(st['step']))))
        if len(steps) > 1:
            raise SnfError(f'parameters carry different Adam step counts {sorted(steps)}: one flat step count is kept')
        self.lr0 = float(group.get('initial_lr', self.lr0))
        self._set_schedule(float(group['lr']), steps.pop() if steps else 0)

    def lr_scheduler_state_dict(self) -> Dict[str, Any]:
        """`ExponentialLR.state_dict()` of sunerf.py:32-33: the schedule is stepped once per batch while lr > 5e-5."""
        self._sync_sched_from_device()
        import math
        k = int(round(math.log(self.lr / self.lr0) / math.log(self.gamma))) if self.lr != self.lr0 else 0
        return {'gamma': self.gamma, 'base_lrs': [self.lr0], 'last_epoch': k, '_step_count': k + 1,
                '_get_lr_called_within_step': False, '_last_lr': [self.lr]}

    def save_checkpoint(self, path: str, extra: Optional[Dict[str, Any]] = None) -> None:
        """A Lightning-layout checkpoint (`state_dict` with the `rendering.` prefix of the reference's modules,
        `optimizer_states`, `lr_schedulers`, `global_step`): the reference's `ckpt_path='last'` resume reads it, and
        `load_checkpoint` restores a RayTrainer from the reference's own `last.ckpt`."""
        sd = OrderedDict(('rendering.' + k, v.detach().cpu().clone()) for k, v in self.r.state_dict().items())
        ck = {'state_dict': sd, 'optimizer_states': [self.optimizer_state_dict()],
              'lr_schedulers': [self.lr_scheduler_state_dict()], 'global_step': self.step_count, 'epoch': 0,
              'pytorch-lightning_version': '1.9.3'}
        ck.update(extra or {})
        torch.save(ck, path)

    def load_checkpoint(self, path: str) -> Dict[str, Any]:
        ck = torch.load(path, map_location='cpu', weights_only=True)
        sd = {k[len('rendering.'):]: v for k, v in ck['state_dict'].items() if k.startswith('rendering.')}
        self.r.load_state_dict(sd, strict=False)                    # sunerf.py:56-59 loads strict=False (in place: copy_)
        self._check_homed()
        if ck.get('optimizer_states'):
            self.load_optimizer_state_dict(ck['optimizer_states'][0])
        for name in ('fine_model', 'coarse_model'):
            getattr(self.r, name)._pack_key = None
        return ck
