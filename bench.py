#!/usr/bin/env python
"""Benchmark of the multislice forward + gradient + update hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU path (oracle port)

Workload (N = 1 and per rank for N > 1): BASELINE config D, `multislice_ptycho_256_theta` -- 256^3 delta/beta object,
72x72 probe, 23x23 = 529 probe positions per angle (step 12, origin -36), 256 slices, far-field detector, lsq loss +
L1 + TV regularisers, Adam (demos/multislice_ptycho_256_theta.py:52-93).  One "step" = one minibatch iteration of
the reference driver with minibatch_size = 529 (all positions of one rotation angle on each rank): rotate the object,
multislice forward, loss, analytic adjoint, transpose rotation, regulariser, [NCCL all-reduce SUM of the object
gradient], fused Adam.  Angles are sharded over ranks (weak scaling: per-GPU work is fixed).  Data are synthetic
(forward-simulated from a seeded phantom with the same kernels before the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'diffraction patterns/sec (fwd+grad+update) at 256^3'
UNIT = 'patterns/s'

CFG_D = dict(obj_size=(256, 256, 256), probe_size=(72, 72), energy_ev=5000., psize_cm=1.e-7,
             alpha_d=1e-9 * 1.7e7, alpha_b=1e-10 * 1.7e7, gamma=1e-9 * 1.7e7, learning_rate=5e-5,
             n_theta=500, probe_mag_sigma=6, probe_phase_sigma=6, probe_phase_max=0.5)


def probe_positions():
    return np.array([(y, x) for y in (np.arange(23) * 12) - 36 for x in (np.arange(23) * 12) - 36], dtype='int64')


def workload_name(mb):
    return ('multislice_ptycho_256_theta (BASELINE configs[3]): 256^3 object, 72x72 probe, 529 positions/angle, 256 slices, '
            'far field, lsq+L1+TV, Adam; minibatch_size=%d (one rotation angle per GPU per step)' % mb)


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clock and throttle reasons DURING the timed region through NVML (nvidia_ml_py) every 20 ms; the
    nvidia-smi query of the recipe buffers its pipe output and returned no rows for sub-second regions."""
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.sm, self.reasons, self.smax, self.power = [], set(), None, []
        self.stop_flag = threading.Event()
        self.th = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(',')[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:      # noqa
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.BITS:
                    if r & bit:
                        self.reasons.add(name)
            except Exception as e:      # noqa
                self.err = repr(e)
                break
            time.sleep(0.02)

    def stop(self):
        if self.th is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable: %s' % self.err], 'samples': 0}
        self.stop_flag.set()
        self.th.join(timeout=1.0)
        # the samples taken while the GPU was busy: the upper half of the power readings marks the loaded interval
        sm = self.sm
        if len(sm) >= 4 and len(self.power) == len(sm):
            thr = 0.5 * (max(self.power) + min(self.power))
            busy = [c for c, p in zip(sm, self.power) if p >= thr]
            if len(busy) >= 2:
                sm = busy
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.smax, 'reasons': sorted(self.reasons),
                'samples': len(self.sm), 'samples_under_load': len(sm), 'power_w_max': max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated (oracle port), all host threads, bounded sample
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sizes=(2, 8, 24), seed=0):
    """Times the oracle's rotate + fwd + autograd + Adam on the config-D geometry on ALL host cores.

    A full 529-pattern step would take minutes on the CPU, so each timed step is a BOUNDED SAMPLE of n patterns of one
    angle, for three sample sizes; the per-step time is linear in n (t = a + b n: a = rotation + regulariser + Adam on
    the 256^3 volume, b = multislice fwd + backward per pattern), and the reported value is the throughput the
    least-squares fit gives for the benchmark's minibatch of 529 patterns, 529 / (a + 529 b) -- an EXTRAPOLATION
    (flagged as such in the JSON line, with the fit residual); the raw timings are returned as well."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core regardless
    from oracle import multislice_oracle as orc
    c = CFG_D
    Y, X, Z = c['obj_size']
    P = list(c['probe_size'])
    rs = np.random.RandomState(seed)
    obj = torch.tensor(np.stack([rs.normal(8.7e-7, 1e-7, (Y, X, Z)), rs.normal(5.1e-8, 1e-8, (Y, X, Z))], -1).astype('float32'))
    pr, pi = orc.gaussian_probe(P, 1, c['probe_mag_sigma'], c['probe_phase_max'], c['probe_phase_sigma'])
    pr = torch.tensor(pr[None].astype('float32'))
    pi = torch.tensor(pi[None].astype('float32'))
    pos = probe_positions()
    voxel_nm = np.array([c['psize_cm']] * 3) * 1e7
    h = orc.get_kernel(voxel_nm[-1], 1240. / c['energy_ev'], voxel_nm, P)
    cfg = dict(this_obj_size=[Y, X, Z], probe_size=P, energy_ev=c['energy_ev'], psize_cm=c['psize_cm'], h=h,
               free_prop_cm='inf', unknown_type='delta_beta')
    reg = dict(alpha_d=c['alpha_d'], alpha_b=c['alpha_b'], gamma=c['gamma'])
    theta_ls = np.linspace(0, 2 * np.pi, c['n_theta'], dtype='float32')
    m = torch.zeros_like(obj)
    v = torch.zeros_like(obj)
    t_of_n = {}
    it = 0
    for n in sizes:
        prj = torch.tensor(rs.uniform(0.5, 1.5, (n, P[0], P[1])).astype('float32'))
        ts = []
        for k in range(warmup + steps):
            t0 = time.perf_counter()
            i_theta = (7 * it + 3) % c['n_theta']
            ind = np.sort(rs.choice(len(pos), n, replace=False))
            lut = orc.rotation_coords([Y, X, Z], theta_ls[i_theta])
            loss, _, _, grads = orc.ptycho_loss_and_grads(obj, pr, pi, pos[ind], prj, lut, cfg, reg=reg, want=('obj',))
            obj, m, v = orc.adam_step(obj, grads['obj'], m, v, it, step_size=c['learning_rate'])
            it += 1
            if k >= warmup:
                ts.append(time.perf_counter() - t0)
        t_of_n[n] = float(np.mean(ts))
    ns = np.array(sorted(t_of_n), dtype='float64')
    ts = np.array([t_of_n[int(n)] for n in ns])
    b, a = np.polyfit(ns, ts, 1)
    a = max(float(a), 0.0)
    b = float(b)
    resid = float(np.max(np.abs(a + b * ns - ts) / ts))
    value = 529.0 / (a + 529.0 * b)
    sample = ('oracle port of the reference PyTorch-CPU path, config-D geometry (256^3, P=72, 256 slices, far field, lsq+L1+TV, '
              'Adam), %d host threads; timed steps of %s patterns of one angle: %s s per step (= %s patterns/s raw); '
              'least-squares fit t = %.2f s + %.3f s/pattern (max residual %.1f %%) EXTRAPOLATED to the 529-pattern minibatch'
              % (torch.get_num_threads(), '/'.join('%d' % n for n in ns), '/'.join('%.2f' % t for t in ts),
                 '/'.join('%.3f' % (n / t) for n, t in zip(ns, ts)), a, b, 100 * resid))
    return {'value': value, 'ms_per_step_529': 1e3 * (a + 529.0 * b), 'cores': torch.get_num_threads(), 'sample': sample,
            'raw': {'patterns': [int(n) for n in ns], 'seconds_per_step': [float(t) for t in ts]}, 'fit_residual': resid}


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
    warmup = 0      # the first (smallest) sample size doubles as the warm-up of the thread pool / allocator
    r = cpu_reference_run(steps, warmup)
    line = {'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup,
            'ms_per_step': r['ms_per_step_529'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': workload_name(args.minibatch), 'obj_size': [256, 256, 256], 'probe_size': [72, 72],
                       'minibatch_size': args.minibatch, 'n_slices': 256},
            'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample'],
                             'extrapolated': True, 'raw': r['raw'], 'fit_residual': r['fit_residual']},
            'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def make_phantom(torch, dev, shape, seed):
    """Seeded smooth 'foam' phantom: delta ~ 0..1e-5, beta = delta / 10."""
    g = torch.Generator(device=dev).manual_seed(seed)
    low = torch.rand((1, 1, 24, 24, 24), device=dev, generator=g)
    up = torch.nn.functional.interpolate(low, size=shape, mode='trilinear', align_corners=False)[0, 0]
    d = torch.clamp(up - 0.45, min=0) * 2e-5
    return torch.stack([d, d / 10], -1).contiguous()


class StepRunner(object):
    """One data-parallel rank of the reference's batch loop (ptychography.py:859-1271), through the public classes
    (PtychographyModel.get_loss_function + Differentiator.get_gradients + AdamOptimizer.apply_gradient)."""

    def __init__(self, torch, ab, dev, rank, world, n_steps_total, mb, dp_mode='nccl', graph=False):
        self.torch, self.ab, self.dev, self.rank, self.world = torch, ab, dev, rank, world
        self.graph = graph
        c = CFG_D
        Y, X, Z = c['obj_size']
        P = list(c['probe_size'])
        self.pos = probe_positions()[:mb]
        self.mb = mb
        full_theta = np.linspace(0, 2 * np.pi, c['n_theta'], dtype='float32')
        # angle schedule: the reference's shuffled order for epoch 0 dealt round-robin to the ranks
        # (ptychography.py:791-847: consecutive chunks of minibatch_size*n_ranks entries, one angle per rank)
        order = np.arange(c['n_theta'])
        np.random.RandomState(0).shuffle(order)
        my_angles = [int(order[(i * world + rank) % c['n_theta']]) for i in range(n_steps_total)]
        self.n_theta = n_steps_total
        self.theta_ls = full_theta[my_angles]          # the model sees exactly the angles this rank will visit
        voxel_nm = np.array([c['psize_cm']] * 3) * 1e7
        h = ab.get_kernel(voxel_nm[-1], 1240. / c['energy_ev'], voxel_nm, P)
        cv = dict(unknown_type='delta_beta', normalize_fft=False, sign_convention=1, rotate_out_of_loop=False,
                  scale_ri_by_k=True, is_minus_logged=False, forward_algorithm='fresnel', stdout_options={},
                  poisson_multiplier=1., common_probe_pos=True, binning=1, prj=None, device_obj=dev, probe_size=P,
                  fresnel_approx=True, two_d_mode=False, minibatch_size=mb, ds_level=1, this_obj_size=[Y, X, Z],
                  energy_ev=c['energy_ev'], psize_cm=c['psize_cm'], h=h, pure_projection=False, n_dp_batch=20,
                  free_prop_cm='inf', optimize_probe_defocusing=False, optimize_probe_pos_offset=False,
                  optimize_prj_pos_offset=False, optimize_all_probe_pos=False, optimize_tilt=False, initial_tilt=None,
                  n_probe_modes=1, n_theta=self.n_theta, precalculate_rotation_coords=False, theta_ls=self.theta_ls,
                  theta_downsample=None, beamstop=None)
        self.fm = ab.PtychographyModel(loss_function_type='lsq', device=dev, common_vars_dict=cv, raw_data_type='magnitude')
        mag, ph = ab.generate_gaussian_map(P, 1, c['probe_mag_sigma'], c['probe_phase_max'], c['probe_phase_sigma'])
        self.probe_real = torch.tensor((mag * np.cos(ph))[None].astype('float32'), device=dev)
        self.probe_imag = torch.tensor((mag * np.sin(ph))[None].astype('float32'), device=dev)
        self.zero_ppc = torch.zeros([self.n_theta, len(self.pos), 2], device=dev)
        self.ind = np.arange(mb)
        # synthetic measured data (adorym/simulation.py:659-671 recipe): predict() with the true object, untimed
        truth = make_phantom(torch, dev, (Y, X, Z), 0)
        frames = []
        with torch.no_grad():
            for i in range(self.n_theta):
                frames.append(self.fm.predict(truth, self.probe_real, self.probe_imag, None, None, i, self.pos, None,
                                              self.zero_ppc, self.ind, None, None))
        del truth
        self.prj_dev = torch.stack(frames)                                   # [n_theta, mb, P, P] device resident
        self.prj_host = self.prj_dev.cpu().pin_memory()                      # host copy for the e2e leg
        self.frame_bytes = int(self.prj_dev[0].numel() * 4)
        rs = np.random.RandomState(0)
        init = np.stack([rs.normal(8.7e-7, 1e-7, (Y, X, Z)), rs.normal(5.1e-8, 1e-8, (Y, X, Z))], -1).astype('float32')
        self.obj0 = torch.tensor(init, device=dev)
        self.fm.add_regularizers([ab.L1Regularizer(c['alpha_d'], c['alpha_b']), ab.TVRegularizer(c['gamma'])])
        # eager differentiator (also the one the per-kernel CUDA-event timing pass uses) and, with --graph, the CUDA-graph
        # replay of loss + gradients (adorym_b200.graph: reconstruct_ptychography(use_cuda_graph=True))
        self.diff_eager = ab.Differentiator()
        self.diff_eager.create_loss_node(self.fm.get_loss_function(), [0])
        self.diff = self.diff_eager
        if graph:
            self.diff_graph = ab.GraphedDifferentiator(self.fm)
            self.diff_graph.create_loss_node(self.fm.get_loss_function(), [0])
            self.diff = self.diff_graph
        # N > 1 on one NVSwitch node: the object and its gradient live in symmetric (peer-mapped) memory and the update is
        # ONE fused kernel per rank (pull-reduce over NVLink + Adam + push); NCCL all-reduce is the fallback
        self.peer = None
        if world > 1 and dp_mode == 'p2p':
            from adorym_b200 import dist as _adist, ops as _ops
            self.peer = _adist.PeerUpdate([Y, X, Z, 2], dev)
            _ops.OBJ_GRAD_OUT = self.peer.gradient_tensor()
        self.reset()

    def reset(self):
        torch, ab = self.torch, self.ab
        if self.peer is not None:
            with torch.no_grad():
                self.peer.object_tensor().copy_(self.obj0)
            self.obj = self.peer.object_tensor().requires_grad_(True)
        elif getattr(self, 'obj', None) is not None:
            with torch.no_grad():           # same storage across resets (a captured graph has adopted it)
                self.obj.detach().copy_(self.obj0)
        else:
            self.obj = self.obj0.clone().requires_grad_(True)
        self.opt = ab.AdamOptimizer('obj', options_dict={'step_size': CFG_D['learning_rate']})
        self.opt.create_container(list(self.obj.shape), False, self.dev)
        self.gradient = torch.zeros_like(self.obj0)
        self.i_step = 0

    def step(self, host_data):
        """One minibatch iteration.  host_data=True: the measured frames live in pinned host memory (H2D copy inside the
        step, via ForwardModel.get_data) and the loss is read back to the host, as reconstruct_ptychography does."""
        torch = self.torch
        i_theta = self.i_step % self.n_theta
        self.fm.prj = self.prj_host if host_data else self.prj_dev
        if host_data:
            # the copy of the NEXT step's frames runs on a side stream under this step's kernels (every step still
            # copies its own frames from pinned host memory: h2d_bytes_per_step is unchanged)
            if self.i_step == 0 or getattr(self.fm, '_prefetched', None) is None:
                self.fm.prefetch_data(i_theta, self.ind)
        args = dict(obj=self.obj, probe_real=self.probe_real, probe_imag=self.probe_imag, probe_defocus_mm=None,
                    probe_pos_offset=None, this_i_theta=i_theta, this_pos_batch=self.pos, prj=self.fm.prj,
                    probe_pos_correction=self.zero_ppc, this_ind_batch=self.ind, tilt_ls=None, prj_pos_offset=None)
        grads = self.diff.get_gradients(**args)
        if host_data:
            self.fm.prefetch_data((self.i_step + 1) % self.n_theta, self.ind)
        # ptychography.py:1065-1067 (`gradient.arr = zeros; gradient.arr += grads[0]` when the buffer is re-initialised,
        # i.e. every step at update_scheme='immediate'): 0 + g is g, so the buffer simply becomes the new gradient
        self.gradient = grads[0]
        from adorym_b200 import ops as _ops, dist as _adist
        slabs = _ops.last_obj_grad_slabs
        with torch.no_grad():
            if self.peer is not None:
                if self.gradient.data_ptr() != self.peer.g.data_ptr():
                    self.peer.gradient_tensor().copy_(self.gradient)
                if slabs is not None and slabs[0].data_ptr() == self.peer.g.data_ptr() and len(slabs[1]) > 1:
                    new = self.peer.update_pipelined(self.opt, self.i_step, slabs[1], **self.opt.options_dict)
                else:
                    new = self.peer.update(self.opt, self.i_step, **self.opt.options_dict)
            elif self.world > 1 and slabs is not None and slabs[0].data_ptr() == self.gradient.data_ptr():
                # gradient all-reduce pipelined by y slabs with the adjoint rotation and with Adam
                new = _adist.allreduce_update_pipelined(self.opt, self.obj, self.gradient, self.i_step, slabs[1],
                                                        **self.opt.options_dict)
            else:
                if self.world > 1:
                    torch.distributed.all_reduce(self.gradient, op=torch.distributed.ReduceOp.SUM)
                new = self.opt.apply_gradient(self.obj, self.gradient, self.i_step, **self.opt.options_dict)
        self.obj = new.requires_grad_(True)
        self.i_step += 1
        return self.fm.current_loss if host_data else None


def _ops_mod():
    from adorym_b200 import ops
    return ops


def dp_check(torch, ab, runner, dev, rank, world):
    """On-hardware check of the data-parallel semantics (untimed, N > 1):
    (1) one DP step's all-reduced object gradient equals the SUM of the single-rank gradients of all ranks' minibatches,
        recomputed serially on rank 0 from the gathered frames (SUM, not mean: adorym/ptychography.py:1113-1114);
    (2) after the timed steps every replica of the object is bit-identical to rank 0's (checked by the caller)."""
    import torch.distributed as dist
    runner.reset()
    i_theta = 0
    runner.fm.prj = runner.prj_dev
    args = dict(obj=runner.obj, probe_real=runner.probe_real, probe_imag=runner.probe_imag, probe_defocus_mm=None,
                probe_pos_offset=None, this_i_theta=i_theta, this_pos_batch=runner.pos, prj=runner.fm.prj,
                probe_pos_correction=runner.zero_ppc, this_ind_batch=runner.ind, tilt_ls=None, prj_pos_offset=None)
    g = runner.diff.get_gradients(**args)[0].clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    frames = [torch.empty_like(runner.prj_dev[0]) for _ in range(world)]
    dist.all_gather(frames, runner.prj_dev[0].contiguous())
    thetas = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(thetas, torch.tensor([float(runner.theta_ls[0])], device=dev))
    out = None
    if rank == 0:
        cv = dict(runner.fm.common_vars)
        cv['theta_ls'] = np.array([float(t.item()) for t in thetas], dtype='float32')
        cv['n_theta'] = world
        fm = ab.PtychographyModel(loss_function_type='lsq', device=dev, common_vars_dict=cv, raw_data_type='magnitude')
        fm.add_regularizers(runner.fm.reg_list)
        fm.prj = torch.stack(frames)
        diff = ab.Differentiator()
        diff.create_loss_node(fm.get_loss_function(), [0])
        zero = torch.zeros([world, len(runner.pos), 2], device=dev)
        gs = None
        for r in range(world):
            a = dict(args, this_i_theta=r, prj=fm.prj, probe_pos_correction=zero)
            gr = diff.get_gradients(**a)[0]
            gs = gr.clone() if gs is None else gs + gr
        out = float((g - gs).double().norm() / gs.double().norm())
    runner.reset()
    p2p = None
    if runner.peer is not None:
        # (3) the fused peer-memory update (adm_p2p_reduce_adam) against all_gather + rank-ordered sum + the single-GPU
        # Adam kernel on rank 0: the same fp32 operations in the same order, so the objects must agree bit for bit
        g = runner.diff.get_gradients(**dict(args, obj=runner.obj))[0]
        gl = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(gl, g.contiguous())
        x0 = runner.obj.detach().clone()
        opt_p = ab.AdamOptimizer('obj', options_dict=dict(runner.opt.options_dict))
        with torch.no_grad():
            if g.data_ptr() != runner.peer.g.data_ptr():
                runner.peer.gradient_tensor().copy_(g)
            slabs = _ops_mod().last_obj_grad_slabs
            if slabs is not None and slabs[0].data_ptr() == runner.peer.g.data_ptr() and len(slabs[1]) > 1:
                # the slab-pipelined form of the update is the one the timed steps use: check THAT one
                xn = runner.peer.update_pipelined(opt_p, 0, slabs[1], **opt_p.options_dict)
            else:
                xn = runner.peer.update(opt_p, 0, **opt_p.options_dict)
        if rank == 0:
            gs = gl[0].clone()
            for r in range(1, world):
                gs += gl[r]
            opt_r = ab.AdamOptimizer('obj', options_dict=dict(runner.opt.options_dict))
            with torch.no_grad():
                xr = opt_r.apply_gradient(x0, gs, 0, **opt_r.options_dict)
            p2p = float((xr - xn.detach()).abs().max().item())
        runner.reset()
    return out, p2p


def run_gpu_arm(args):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU arm)')
    from adorym_b200 import dist as adist
    rank, world, local = adist.init_from_env('nccl')
    if world != args.gpus and world > 1:
        args.gpus = world
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    import adorym_b200 as ab
    from adorym_b200 import _lib
    mb = args.minibatch
    K, W = args.steps, max(args.warmup, 3)
    if args.grad_slabs is None:
        args.grad_slabs = 1 if args.dp == 'p2p' else 2
    from adorym_b200 import ops as _ops
    _ops.GRAD_SLABS = args.grad_slabs if world > 1 else 1
    runner = StepRunner(torch, ab, dev, rank, world, K + W, mb, dp_mode=args.dp, graph=args.graph)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(host_data, with_kernel_timing):
        # the per-kernel CUDA-event pass needs eager launches; everything else replays the CUDA graph when --graph is given
        runner.diff = runner.diff_eager if (with_kernel_timing or not runner.graph) else runner.diff_graph
        runner.reset()
        for _ in range(W):
            runner.step(host_data)
        barrier()
        if with_kernel_timing:
            _lib.enable_timing(True)
        l0 = _lib.launch_count
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            runner.step(host_data)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count - l0
        ksum = _lib.timing_summary() if with_kernel_timing else {}
        _lib.enable_timing(False)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, ksum

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_eager = None
    if runner.graph:
        ms_dev, _, _ = timed(False, False)                 # value: CUDA-graph replay of loss + gradients, Adam eager
        clocks = sampler.stop() if rank == 0 else None
        ms_eager, launches, ksum = timed(False, True)      # the same steps launched eagerly, a CUDA-event pair per kernel
    else:
        ms_dev, launches, ksum = timed(False, True)
        clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(True, False)
    runner.diff = runner.diff_eager
    frame_bytes = runner.frame_bytes
    dp = None
    if world > 1:
        # replicas after the timed steps: every rank's object must equal rank 0's bit for bit
        ref = runner.obj.detach().clone()
        torch.distributed.broadcast(ref, src=0)
        d = (ref - runner.obj.detach()).abs().max().reshape(1)
        torch.distributed.all_reduce(d, op=torch.distributed.ReduceOp.MAX)
        sum_err, p2p_err = dp_check(torch, ab, runner, dev, rank, world)
        dp = {'max_abs_replica_difference': float(d.item()), 'allreduced_vs_summed_single_rank_gradient_rel': sum_err,
              'p2p_update_vs_allgather_sum_adam_max_abs': p2p_err,
              'grad_slabs': _ops.GRAD_SLABS, 'update': 'p2p fused pull-reduce + Adam + push (adm_p2p_reduce_adam, %s)'
              % (('NVLS multimem.ld_reduce / multimem.st' if runner.peer is not None and runner.peer.mc_x is not None
                  else 'peer loads / stores') + ('; pipelined by %d y slabs with the transpose chain' % _ops.GRAD_SLABS
                                                 if _ops.GRAD_SLABS > 1 else ''))
              if runner.peer is not None else 'NCCL all-reduce (slab pipelined) + Adam'}

    # ---- secondary figure: the reference demo's own schedule (minibatch_size = 2, update after every minibatch,
    # demos/multislice_ptycho_256_theta.py:63,93); every step still rotates and Adam-updates the whole 256^3 volume, so
    # it is bounded by those passes and by the 256 sequential slices of only two wavefields ----
    ref_sched = None
    if args.minibatch != 2 and not args.no_reference_schedule:
        del runner
        torch.cuda.empty_cache()
        K2 = min(K, 20)
        _ops.OBJ_GRAD_OUT = None
        runner = StepRunner(torch, ab, dev, rank, world, K2 + W, 2)
        mbk, Kk = mb, K
        mb, K = 2, K2
        ms2, _, _ = timed(False, False)
        mb, K = mbk, Kk
        ref_sched = {'minibatch_size': 2, 'update_scheme': 'immediate', 'steps': K2, 'ms_per_step': ms2 / K2,
                     'value': world * 2 * K2 / (ms2 * 1e-3), 'unit': UNIT}
        runner = None

    if rank != 0:
        return
    value = world * mb * K / (ms_dev * 1e-3)
    e2e = world * mb * K / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (algorithmic bytes) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    P, Nz = 72, 256
    p8 = P * P * Nz * 8
    d8 = P * P * 8
    d4 = P * P * 4
    V8 = 256 ** 3 * 8
    # ALGORITHMIC bytes per launch (minimum HBM traffic, wavefield shared-memory resident):
    #   adjoint : re-read (delta, beta) + accumulate the gradient = 2 p8 per pattern (+ the detector residual)
    #   forward : read (delta, beta) + write the detector field    = p8 + d8 per pattern
    # The per-slice psi checkpoints (2 p8 per pattern: written by the forward, read by the adjoint) are DESIGN traffic,
    # not algorithmic; the figure that includes them is reported separately as `frac_with_checkpoints`.
    alg = {'adm_multislice_bwd': mb * (2 * p8 + d8), 'adm_multislice_fwd': mb * (p8 + d8)}
    ckpt_extra = mb * p8
    dom = max(alg.keys(), key=lambda k: ksum.get(k, (0, 0.0))[1])
    n_dom, t_dom = ksum.get(dom, (1, float('nan')))
    t_launch = t_dom / max(n_dom, 1) * 1e-3
    achieved = alg[dom] / t_launch / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(dom)
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg[dom],
                'avg_launch_ms': t_dom / max(n_dom, 1),
                'frac_with_checkpoints': (alg[dom] + ckpt_extra) / t_launch / 1e9 / peak,
                'dram_frac': (traffic / t_launch / 1e9 / peak) if traffic else None,
                'kernel_ms_per_step': {k: v[1] / K for k, v in sorted(ksum.items())},
                'note': 'the multislice kernels are bound by instruction throughput (FP32 / issue / shared-memory pipes of the '
                        'FFT passes), not by HBM: see DESIGN.md'}
    if ms_eager is not None:
        roofline['kernel_timing_pass'] = ('the same %d steps launched eagerly with a CUDA-event pair around every kernel '
                                          '(%.3f ms per step); value / e2e replay loss + gradients as one CUDA graph' % (K, ms_eager / K))
    # whole-step algorithmic bytes: B (3 p8 + d4) + 4 V8 (rotation fwd + adj) + 7 V8 (Adam)
    step_bytes = mb * (3 * p8 + d4) + 4 * V8 + 7 * V8
    roofline['step_frac'] = (step_bytes / (ms_dev / K * 1e-3) / 1e9) / peak
    roofline['step_algorithmic_bytes'] = step_bytes
    roofline['step_frac_with_checkpoints'] = ((step_bytes + 2 * mb * p8) / (ms_dev / K * 1e-3) / 1e9) / peak
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(1, 0)
        cpu = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample'],
               'extrapolated': True, 'raw': r['raw'], 'fit_residual': r['fit_residual']}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms_dev / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': workload_name(mb), 'obj_size': [256, 256, 256], 'probe_size': [72, 72],
                       'minibatch_size': mb, 'n_slices': 256,
                       'execution': ('loss + gradients of a step replayed as ONE CUDA graph (adorym_b200.graph), optimizer '
                                     'update eager' if args.graph else 'eager launches'),
                       'parallelism': 'dp%d (angles sharded; object gradient summed over ranks %s)'
                                      % (world, 'by the fused NVLink peer-memory update kernel' if (world > 1 and args.dp == 'p2p')
                                         else 'with an NCCL all-reduce'),
                       'l2_policy': 'inputs larger than L2: 134 MB volumes, %.1f GB psi checkpoints per step vs 126 MB L2'
                                    % (mb * p8 / 1e9)},
            'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks,
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': frame_bytes + mb * 8,
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / K},
            'gpu_launches': launches, 'reference_default_schedule': ref_sched, 'dp_check': dp}
    print(json.dumps(line), flush=True)


def run_other_configs(args):
    """The other named shapes of BASELINE.json (configs A, B, C, E), with the same clocks sampler as the headline line: one
    JSON line per configuration and execution mode (eager / CUDA graph).  Under torchrun (N > 1) the ranks run data parallel
    as the driver does -- every rank its own minibatch (angle) of the global chunk, gradients summed with NCCL, identical
    update; weak scaling -- except config E, which has nothing to shard.  Secondary figures: `python bench.py` without
    --config is the driver's contract line on config D."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --config: no CUDA device (there is no CPU fallback)')
    from adorym_b200 import dist as adist
    rank, world, local = adist.init_from_env('nccl')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import bench_configs as bc
    bc.DEV, bc.RANK, bc.WORLD = dev, rank, world
    for cfg in args.config:
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        lines = bc.run([cfg], steps=args.steps, warmup=max(args.warmup, 3))
        clocks = sampler.stop() if rank == 0 else None
        if rank != 0:
            continue
        for l in lines:
            out = {'metric': METRIC.replace('at 256^3', 'config ' + l['config']), 'value': l['value'], 'unit': UNIT, 'n_gpus': world,
                   'steps': l['steps'], 'warmup': l['warmup'], 'ms_per_step': l['ms_per_step'], 'higher_is_better': True,
                   'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                   'config': {'workload': l['workload'] + (' [loss + gradients replayed as one CUDA graph]' if l.get('cuda_graph') else ' [eager launches]'),
                              'cuda_graph': bool(l.get('cuda_graph')), 'patterns_per_step': l['patterns_per_step'],
                              'parallelism': 'dp%d: every rank its own minibatch of the global chunk, all gradients SUMMED with one '
                                             'flat NCCL all-reduce, identical update' % world,
                              'l2_policy': 'working set is L2 resident by construction (<= 8 MB): launch / latency bound'},
                   'roofline': {'bound': 'hbm', 'achieved': l['algorithmic_bytes_per_step'] / (l['ms_per_step'] * 1e-3) / 1e9,
                                'peak': l['peak_gbs'], 'unit': 'GB/s', 'frac': l['hbm_roofline_frac'], 'traffic': None},
                   'clocks': clocks}
            print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--minibatch', type=int, default=529)
    ap.add_argument('--dp', default='p2p', choices=['p2p', 'nccl'],
                    help='N > 1: fused peer-memory update kernel (default) or NCCL all-reduce + Adam')
    ap.add_argument('--grad-slabs', type=int, default=None,
                    help='N > 1: y slabs by which the gradient exchange is pipelined with its production (default: 1 = not '
                         'pipelined for the fused peer-memory update -- measured faster, DESIGN.md section 6 -- and 2 for NCCL)')
    ap.add_argument('--config', nargs='*', default=None, choices=['A', 'B', 'C', 'E'],
                    help='secondary lines for the other named shapes (default: the contract line on config D)')
    ap.add_argument('--graph', action='store_true',
                    help='replay loss + gradients of a step as one CUDA graph (adorym_b200.graph) in the value and e2e legs')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-reference-schedule', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    elif args.config:
        run_other_configs(args)
    else:
        run_gpu_arm(args)
    try:
        import torch.distributed as d
        if d.is_available() and d.is_initialized():
            d.destroy_process_group()
    except Exception:
        pass


if __name__ == '__main__':
    main()
