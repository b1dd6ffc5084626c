#!/usr/bin/env python
"""Secondary throughput figures on one B200 for the OTHER named shapes of BASELINE.json (configs A, B, C, E; config D is
bench.py): forward + gradient(s) + Adam per minibatch through the public classes, synthetic data of the named shape,
CUDA-event timing after warm-up.  Prints one JSON line per config with patterns/s, ms per step and the fraction of the
HBM roofline for the algorithmic bytes of SURVEY section 8d (small by construction: these working sets are L2 resident
and the steps are launch / latency bound).

    python tools/bench_configs.py [A B C E]
"""
import json
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adorym_b200 as ab  # noqa: E402

DEV = torch.device('cuda', 0)
RANK, WORLD = 0, 1          # set by bench.py --config under torchrun: data parallelism as in the driver (each rank its own
                            # minibatch of the global chunk, gradients SUMMED over ranks with one NCCL all-reduce, identical update)
PEAK = 6531.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def base_cv(**kw):
    cv = dict(unknown_type='delta_beta', normalize_fft=False, sign_convention=1, rotate_out_of_loop=False,
              scale_ri_by_k=True, is_minus_logged=False, forward_algorithm='fresnel', stdout_options={},
              poisson_multiplier=1., common_probe_pos=True, binning=1, prj=None, device_obj=DEV, fresnel_approx=True,
              two_d_mode=False, ds_level=1, pure_projection=False, n_dp_batch=20, optimize_probe_defocusing=False,
              optimize_probe_pos_offset=False, optimize_prj_pos_offset=False, optimize_all_probe_pos=False,
              optimize_tilt=False, initial_tilt=None, n_probe_modes=1, precalculate_rotation_coords=False,
              theta_downsample=None, beamstop=None)
    cv.update(kw)
    return cv


GRAPH = False     # True: the loss + gradient evaluation replays a CUDA graph (adorym_b200.graph.GraphedDifferentiator)
STEPS = None      # overridden by bench.py --config (steps, warmup)
WARMUP = None
LINES = []


def sum_over_ranks(grads):
    """ptychography.py:1113-1114 / optimizers.py:1025-1079: SUM (not mean) of every gradient over the ranks -- one flat
    NCCL all-reduce for all of them."""
    if WORLD == 1:
        return list(grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
    out, o = [], 0
    for g in grads:
        out.append(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    return out


def time_steps(step, n_warm=5, n=30):
    n_warm = WARMUP if WARMUP is not None else n_warm
    n = STEPS if STEPS is not None else n
    for _ in range(n_warm):
        step()
    if WORLD > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    if WORLD > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    time_steps.last = (n, n_warm)
    ms = e0.elapsed_time(e1) / n
    if WORLD > 1:                     # the job's step time is the slowest rank's
        t = torch.tensor([ms], device=DEV)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def report(name, workload, patterns_per_step, ms, alg_bytes):
    n, n_warm = getattr(time_steps, 'last', (None, None))
    # weak scaling: every rank processes `patterns_per_step` patterns per step; the roofline fraction is per GPU
    line = {'config': name, 'cuda_graph': bool(GRAPH), 'workload': workload, 'patterns_per_step': patterns_per_step * WORLD,
            'ms_per_step': ms, 'n_gpus': WORLD,
            'value': patterns_per_step * WORLD / (ms * 1e-3), 'unit': 'patterns/s', 'steps': n, 'warmup': n_warm,
            'algorithmic_bytes_per_step': alg_bytes,
            'hbm_roofline_frac': alg_bytes / (ms * 1e-3) / 1e9 / PEAK, 'peak_gbs': PEAK}
    LINES.append(line)
    if STEPS is None and RANK == 0:
        print(json.dumps(line), flush=True)


def ptycho_runner(obj_size, P, pos, M, unknown_type, energy_ev, psize_cm, raw_data_type, mb, refine, lr_obj, seed=0):
    """PtychographyModel step: gradients w.r.t. object (+ probe, + all probe positions), Adam on each."""
    Y, X, Z = obj_size
    rs = np.random.RandomState(seed)
    voxel_nm = np.array([psize_cm] * 3) * 1e7
    h = ab.get_kernel(voxel_nm[-1], 1240. / energy_ev, voxel_nm, [P, P])
    n_pos = len(pos)
    cv = base_cv(unknown_type=unknown_type, probe_size=[P, P], two_d_mode=(Z == 1), minibatch_size=mb,
                 this_obj_size=[Y, X, Z], energy_ev=energy_ev, psize_cm=psize_cm, h=h, free_prop_cm='inf',
                 optimize_all_probe_pos=refine, n_probe_modes=M, n_theta=1, theta_ls=np.zeros(1, 'float32'))
    fm = ab.PtychographyModel(loss_function_type='lsq', device=DEV, common_vars_dict=cv, raw_data_type=raw_data_type)
    mag, ph = ab.generate_gaussian_map([P, P], 1, 6, 0.5, 6)
    pr0, pi0 = mag * np.cos(ph), mag * np.sin(ph)
    pr = np.stack([pr0 * (1 + 0.2 * m) + 0.02 * rs.normal(size=(P, P)) * np.abs(pr0) for m in range(M)]).astype('float32')
    pi = np.stack([pi0 * (1 - 0.1 * m) + 0.02 * rs.normal(size=(P, P)) * np.abs(pr0) for m in range(M)]).astype('float32')
    probe_real = torch.tensor(pr, device=DEV, requires_grad=refine)
    probe_imag = torch.tensor(pi, device=DEV, requires_grad=refine)
    if unknown_type == 'delta_beta':
        truth = np.stack([rs.uniform(5e-5, 0.039, (Y, X, Z)), np.zeros((Y, X, Z))], -1).astype('float32')
        init = np.stack([np.full((Y, X, Z), 1e-3), np.zeros((Y, X, Z))], -1).astype('float32')
    else:
        phs = rs.uniform(-0.5, 0.5, (Y, X, Z))
        truth = np.stack([np.cos(phs), np.sin(phs)], -1).astype('float32')
        init = np.stack([np.ones((Y, X, Z)), np.zeros((Y, X, Z))], -1).astype('float32')
    ppc = torch.zeros([1, n_pos, 2], device=DEV, requires_grad=refine)
    n_batches = n_pos // mb
    # synthetic data: predict with the true object (adorym/simulation.py:659-671 recipe)
    frames = []
    with torch.no_grad():
        t = torch.tensor(truth, device=DEV)
        for i in range(n_batches):
            ind = np.arange(i * mb, (i + 1) * mb)
            a = fm.predict(t, probe_real.detach(), probe_imag.detach(), None, None, 0, pos[ind], None, ppc.detach(), ind,
                           None, None)
            frames.append(a ** 2 if raw_data_type == 'intensity' else a)
    prj = torch.zeros([1, n_pos, P, P], device=DEV)
    prj[0, :n_batches * mb] = torch.cat(frames)
    fm.prj = prj
    diff = ab.GraphedDifferentiator(fm) if GRAPH else ab.Differentiator()
    opt_args = [0]
    if refine:
        opt_args += [fm.get_argument_index('probe_real'), fm.get_argument_index('probe_imag'),
                     fm.get_argument_index('probe_pos_correction')]
    diff.create_loss_node(fm.get_loss_function(), opt_args)
    state = {'obj': torch.tensor(init, device=DEV, requires_grad=True), 'i': 0}
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': lr_obj})
    opt.create_container([Y, X, Z, 2], False, DEV)
    opt_p = ab.AdamOptimizer('probe', options_dict={'step_size': 1e-3})
    opt_p.create_container([M, P, P, 2], False, DEV)
    probe = {'t': torch.stack([probe_real.detach(), probe_imag.detach()], -1).contiguous()}

    n_chunks = max(1, n_batches // WORLD)          # global chunks of WORLD minibatches; rank r takes the r-th (ptychography.py:905-907)

    def step():
        i = ((state['i'] % n_chunks) * WORLD + RANK) % n_batches
        ind = np.arange(i * mb, (i + 1) * mb)
        p_r = probe['t'][..., 0].contiguous().requires_grad_(refine)
        p_i = probe['t'][..., 1].contiguous().requires_grad_(refine)
        grads = diff.get_gradients(obj=state['obj'], probe_real=p_r, probe_imag=p_i, probe_defocus_mm=None,
                                   probe_pos_offset=None, this_i_theta=0, this_pos_batch=pos[ind], prj=prj,
                                   probe_pos_correction=ppc, this_ind_batch=ind, tilt_ls=None, prj_pos_offset=None)
        grads = sum_over_ranks(grads)
        with torch.no_grad():
            new = opt.apply_gradient(state['obj'], grads[0], state['i'], **opt.options_dict)
            if refine:
                g = torch.stack([grads[1], grads[2]], -1).contiguous()
                probe['t'] = opt_p.apply_gradient(probe['t'], g, state['i'], **opt_p.options_dict)
        state['obj'] = new.requires_grad_(True)
        state['i'] += 1

    return step


def config_A():
    """tests/test_multislice_tomography_64.py shape: 64^3 full-field multislice tomography, one projection per step."""
    Y = X = Z = 64
    n_theta = 50
    energy_ev, psize_cm = 800., 0.67e-7
    voxel_nm = np.array([psize_cm] * 3) * 1e7
    h = ab.get_kernel(voxel_nm[-1], 1240. / energy_ev, voxel_nm, [Y, X])
    theta_ls = np.linspace(0, 2 * np.pi, n_theta, dtype='float32')
    cv = base_cv(probe_size=[Y, X], minibatch_size=1, this_obj_size=[Y, X, Z], energy_ev=energy_ev, psize_cm=psize_cm, h=h,
                 free_prop_cm=0, n_theta=n_theta, theta_ls=theta_ls)
    fm = ab.SingleBatchFullfieldModel(loss_function_type='lsq', device=DEV, common_vars_dict=cv, raw_data_type='magnitude')
    rs = np.random.RandomState(0)
    truth = torch.tensor(np.stack([rs.uniform(0, 2e-3, (Y, X, Z)), rs.uniform(0, 3e-4, (Y, X, Z))], -1).astype('float32'),
                         device=DEV)
    pr = torch.ones((1, Y, X), device=DEV)
    pi = torch.zeros((1, Y, X), device=DEV)
    pos = np.array([[0, 0]])
    ppc = torch.zeros([n_theta, 1, 2], device=DEV)
    with torch.no_grad():
        prj = torch.stack([fm.predict(truth, pr, pi, None, None, i, pos, None, ppc, np.array([0]), None, None)
                           for i in range(n_theta)])
    fm.prj = prj
    fm.add_regularizers([ab.L1Regularizer(1e-9 * 64 ** 3, 1e-10 * 64 ** 3)])
    diff = ab.GraphedDifferentiator(fm) if GRAPH else ab.Differentiator()
    diff.create_loss_node(fm.get_loss_function(), [0])
    state = {'obj': torch.zeros((Y, X, Z, 2), device=DEV).add_(1e-6).requires_grad_(True), 'i': 0}
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': 1e-7})
    opt.create_container([Y, X, Z, 2], False, DEV)

    def step():
        it = (7 * (state['i'] * WORLD + RANK) + 3) % n_theta          # every rank its own angle of the global chunk
        grads = diff.get_gradients(obj=state['obj'], probe_real=pr, probe_imag=pi, probe_defocus_mm=None,
                                   probe_pos_offset=None, this_i_theta=it, this_pos_batch=pos, prj=prj,
                                   probe_pos_correction=ppc, this_ind_batch=np.array([0]), tilt_ls=None,
                                   prj_pos_offset=None)
        grads = sum_over_ranks(grads)
        with torch.no_grad():
            new = opt.apply_gradient(state['obj'], grads[0], state['i'], **opt.options_dict)
        state['obj'] = new.requires_grad_(True)
        state['i'] += 1

    ms = time_steps(step)
    V8 = Y * X * Z * 8
    report('A', '64^3 full-field multislice tomography (tests/test_multislice_tomography_64.py shape), 1 projection per step, '
           'rotation + 64 slices + L1 + Adam', 1, ms, 3 * V8 + 4 * V8 + 7 * V8)


def config_B():
    pos = np.array([(y, x) for y in np.arange(-10, 246, 5) for x in np.arange(-10, 246, 5)], dtype='int64')
    for mb in (52, 2704):
        step = ptycho_runner((256, 256, 1), 72, pos, 5, 'delta_beta', 5000., 1e-7, 'magnitude', mb, True, 4e-3)
        ms = time_steps(step, n=20)
        p8 = 72 * 72 * 8
        report('B', '2-D ptychography 256^2 x 1, P = 72, 2704 positions, 5 probe modes, probe + position refinement, '
               'minibatch %d' % mb, mb, ms, mb * (3 * p8 + 72 * 72 * 4) + 7 * 256 * 256 * 8)


def config_C():
    ys = np.arange(-20, 618 - 44, 8)
    xs = np.arange(-20, 606 - 44, 8)
    rs = np.random.RandomState(0)
    pos = np.array([(y, x) for y in ys for x in xs], dtype='int64') + rs.randint(-1, 2, size=(len(ys) * len(xs), 2))
    pos = pos[: (len(pos) // 35) * 35]
    step = ptycho_runner((618, 606, 1), 64, pos, 5, 'real_imag', 8801.12, 1.3279e-6, 'intensity', 35, True, 1e-3)
    ms = time_steps(step, n=30)
    p8 = 64 * 64 * 8
    report('C', '2-D fly-scan ptychography 618 x 606 x 1 (siemens_star_aps_2idd shape), P = 64, %d positions, 5 modes, '
           'minibatch 35, probe + all-position refinement' % len(pos), 35, ms,
           35 * (3 * p8 + 64 * 64 * 4) + 7 * 618 * 606 * 8)


def config_E():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from cases import make_md_inputs, MD_CASES as M
    from test_gpu_multidist import md_common_vars
    N = 1024
    M['_bench'] = dict(M['multidist256_shift'], N=N, seed=77)
    inp = make_md_inputs('_bench')
    c = inp['cfg']
    D = len(c['dists'])
    prj_full = torch.tensor(inp['prj'][None], device=DEV)
    fm = ab.MultiDistModel(loss_function_type=c['loss'], distribution_mode=None, device=DEV,
                           common_vars_dict=md_common_vars(c, DEV, prj_full, True), raw_data_type=c['raw_data_type'])
    pr = torch.tensor(inp['probe_real'], device=DEV)
    pi = torch.tensor(inp['probe_imag'], device=DEV)
    state = {'obj': torch.tensor(inp['obj'], device=DEV, requires_grad=True), 'i': 0,
             'ppc': torch.tensor(inp['probe_pos_correction'], device=DEV, requires_grad=True)}
    diff = ab.GraphedDifferentiator(fm) if GRAPH else ab.Differentiator()
    diff.create_loss_node(fm.get_loss_function(), [0, fm.get_argument_index('probe_pos_correction')])
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': 1e-2})
    opt.create_container([N, N, 1, 2], False, DEV)

    def step():
        grads = diff.get_gradients(obj=state['obj'], probe_real=pr, probe_imag=pi, probe_defocus_mm=None,
                                   probe_pos_offset=None, this_i_theta=0, this_pos_batch=np.array([[0, 0]]), prj=prj_full,
                                   probe_pos_correction=state['ppc'], this_ind_batch=np.array([0]),
                                   free_prop_cm=c['dists'], safe_zone_width=0, prj_affine_ls=None, ctf_lg_kappa=None,
                                   prj_pos_offset=None)
        with torch.no_grad():
            new = opt.apply_gradient(state['obj'], grads[0], state['i'], **opt.options_dict)
        state['obj'] = new.requires_grad_(True)
        state['i'] += 1

    ms = time_steps(step, n=30)
    report('E', '2-D multi-distance holography, 4 distances, 1024^2 frames, data-shift refinement (4 patterns per iteration)',
           D, ms, 92e6)


def run(which, steps=None, warmup=None, graph=(False, True)):
    """-> list of result dicts (bench.py --config adds the clocks record and prints them); every configuration is timed
    eagerly and with the CUDA-graph replay of the loss + gradient evaluation."""
    global STEPS, WARMUP, GRAPH
    STEPS, WARMUP = steps, warmup
    del LINES[:]
    for w in which:
        if w == 'E' and WORLD > 1:
            continue          # one object, one set of holograms per iteration: nothing to shard (replicas only)
        for g in graph:
            GRAPH = g
            {'A': config_A, 'B': config_B, 'C': config_C, 'E': config_E}[w]()
    GRAPH = False
    return list(LINES)


if __name__ == '__main__':
    run(sys.argv[1:] or ['A', 'B', 'C', 'E'])
