"""GPU parity tests (run with `-m gpu` on the B200 box): the sm_100a kernels, called through the C ABI via the
reference-shaped Python API, against (i) the oracle evaluated in fp64 on the same seeded inputs and (ii) the golden
vectors produced by the unmodified reference.  Tolerance (north_star): 1e-4 relative on per-pattern loss and
gradients in fp32 for one iteration (1e-3 on the gradient of the ill-conditioned Poisson case, see
test_oracle_golden.py)."""
import os
import numpy as np
import pytest
import torch

from cases import CASES, make_inputs
from oracle_util import case_cfg, run_oracle_case, rel_err, compact_grad
from oracle import multislice_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-4


def _common_vars(inp, cfg, dev, prj_full, theta_ls, n_theta):
    c = inp['cfg']
    Y, X, Z = c['obj_size']
    return dict(unknown_type=c['unknown_type'], normalize_fft=c.get('normalize_fft', False),
                sign_convention=c.get('sign_convention', 1), rotate_out_of_loop=c.get('rotate_out_of_loop', False),
                scale_ri_by_k=True,
                is_minus_logged=False, forward_algorithm='fresnel', stdout_options={},
                poisson_multiplier=c.get('poisson_multiplier', 1.), common_probe_pos=True, binning=c.get('binning', 1),
                prj=prj_full,
                device_obj=dev, probe_size=list(c['probe_size']), fresnel_approx=True, two_d_mode=(Z == 1),
                minibatch_size=len(inp['pos']), ds_level=1, this_obj_size=[Y, X, Z], energy_ev=c['energy_ev'],
                psize_cm=c['psize_cm'], h=cfg['h'], pure_projection=c.get('pure_projection', False), n_dp_batch=20,
                free_prop_cm=c['free_prop_cm'],
                optimize_probe_defocusing=False, optimize_probe_pos_offset=c.get('optimize_probe_pos_offset', False),
                optimize_prj_pos_offset=c.get('optimize_prj_pos_offset', False),
                optimize_all_probe_pos=c.get('optimize_all_probe_pos', False), optimize_tilt=c.get('optimize_tilt', False),
                initial_tilt=None,
                n_probe_modes=c['n_modes'], n_theta=n_theta, precalculate_rotation_coords=False, theta_ls=theta_ls,
                theta_downsample=None, beamstop=inp.get('beamstop'))


def run_gpu_case(name):
    """Drive the case exactly as the reference driver would: model.get_loss_function() + Differentiator."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    inp, cfg, coord, reg = case_cfg(name)
    c = inp['cfg']
    n_theta = c['n_theta']
    P = c['probe_size']
    theta_ls = np.linspace(0, 2 * np.pi, n_theta, dtype='float32')
    theta_ls[c['i_theta']] = c['theta']
    prj_full = np.zeros([n_theta, 16, P[0], P[1]], dtype='float32')
    prj_full[c['i_theta'], inp['this_ind_batch']] = inp['prj']
    cv = _common_vars(inp, cfg, dev, prj_full, theta_ls, n_theta)
    Model = {'fullfield': ab.SingleBatchFullfieldModel, 'sparse': ab.SparseMultisliceModel}.get(c['kind'], ab.PtychographyModel)
    fm = Model(loss_function_type=c['loss'], distribution_mode=None, device=dev, common_vars_dict=cv,
               raw_data_type=c['raw_data_type'])
    regs = []
    if c.get('alpha_d') and c.get('reweighted', False):
        r = ab.ReweightedL1Regularizer(c['alpha_d'], c['alpha_b'], unknown_type=c['unknown_type'])
        r.update_l1_weight(torch.tensor(inp['weight_l1'], device=dev))
        regs.append(r)
    elif c.get('alpha_d'):
        regs.append(ab.L1Regularizer(c['alpha_d'], c['alpha_b'], unknown_type=c['unknown_type']))
    if c.get('gamma'):
        regs.append(ab.TVRegularizer(c['gamma'], unknown_type=c['unknown_type']))
    fm.add_regularizers(regs)
    t = lambda a, rg=True: torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=rg)
    obj, pr, pi = t(inp['obj']), t(inp['probe_real']), t(inp['probe_imag'])
    has_ppc = inp['probe_pos_correction'] is not None
    ppc = t(inp['probe_pos_correction']) if has_ppc else t(np.zeros([n_theta, 16, 2]), False)
    args = dict(obj=obj, probe_real=pr, probe_imag=pi, probe_defocus_mm=torch.tensor(0.0, device=dev),
                probe_pos_offset=torch.zeros([n_theta, 2], device=dev), this_i_theta=c['i_theta'],
                this_pos_batch=inp['pos'], prj=prj_full, probe_pos_correction=ppc, this_ind_batch=inp['this_ind_batch'],
                tilt_ls=np.zeros([3, n_theta]), prj_pos_offset=torch.zeros([n_theta, 2], device=dev))
    if c['kind'] == 'sparse':
        del args['tilt_ls']
        args['slice_pos_cm_ls'] = np.array(c['slice_pos_cm'])
        if c.get('optimize_slice_pos', False):
            cv['optimize_slice_pos'] = True
            args['slice_pos_cm_ls'] = t(np.array(c['slice_pos_cm']))
        args = {k: args[k] for k in fm.argument_ls}          # the Differentiator indexes the kwargs BY POSITION
    opt_args, names = [0], ['obj']
    if c.get('optimize_probe', False):
        opt_args += [fm.get_argument_index('probe_real'), fm.get_argument_index('probe_imag')]
        names += ['probe_real', 'probe_imag']
    if c.get('optimize_probe_pos_offset', False):
        off = np.zeros([n_theta, 2], 'float32')
        off[c['i_theta']] = c['probe_pos_offset']
        args['probe_pos_offset'] = t(off)
        opt_args += [fm.get_argument_index('probe_pos_offset')]
        names += ['probe_pos_offset']
    if c.get('optimize_prj_pos_offset', False):
        off = np.zeros([n_theta, 2], 'float32')
        off[c['i_theta']] = c['prj_pos_offset']
        args['prj_pos_offset'] = t(off)
        opt_args += [fm.get_argument_index('prj_pos_offset')]
        names += ['prj_pos_offset']
    if has_ppc:
        opt_args += [fm.get_argument_index('probe_pos_correction')]
        names += ['probe_pos_correction']
    if c.get('optimize_slice_pos', False):
        opt_args += [fm.get_argument_index('slice_pos_cm_ls')]
        names += ['slice_pos_cm_ls']
    if c.get('optimize_tilt', False):
        tilt = np.zeros([3, n_theta], 'float32')
        tilt[0] = theta_ls
        tilt[:, c['i_theta']] = c['tilt']
        args['tilt_ls'] = t(tilt)
        opt_args += [fm.get_argument_index('tilt_ls')]
        names += ['tilt_ls']
    diff = ab.Differentiator()
    diff.create_loss_node(fm.get_loss_function(), opt_args)
    grads = diff.get_gradients(**args)
    loss = fm.current_loss
    with torch.no_grad():
        pred = fm.predict(**args)
        if isinstance(pred, tuple):
            pred = torch.sqrt(pred[0] ** 2 + pred[1] ** 2)
    torch.cuda.synchronize()
    return inp, loss, pred.cpu().numpy(), {n: g.cpu().numpy() for n, g in zip(names, grads)}, fm


@pytest.mark.parametrize('name', sorted(CASES.keys()))
def test_loss_and_gradients_match_oracle_and_reference(name):
    inp, loss, pred, grads, fm = run_gpu_case(name)
    _, loss64, pred64, grads64 = run_oracle_case(name, torch.float64)
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    gtol = 1e-3 if 'poisson' in name else TOL
    # loss (whole minibatch) and per-pattern magnitudes
    assert abs(loss - loss64) <= TOL * abs(loss64)
    assert abs(loss - float(gold['loss'])) <= TOL * abs(float(gold['loss']))
    for b in range(pred.shape[0]):
        assert rel_err(pred[b], pred64[b]) < TOL, 'pattern %d' % b
    assert rel_err(pred, gold['pred']) < TOL
    # object gradient: full array vs fp64 oracle, compact form vs the reference golden
    assert rel_err(grads['obj'], grads64['obj']) < gtol
    cg = compact_grad(grads['obj'])
    assert abs(cg['norm'] - float(gold['gobj_norm'])) <= gtol * float(gold['gobj_norm'])
    for k in ('proj0', 'proj1', 'proj2', 'sample'):
        assert rel_err(cg[k], gold['gobj_' + k]) < gtol, k
    for k in ('probe_real', 'probe_imag', 'probe_pos_correction'):
        if k in grads:
            assert rel_err(grads[k], grads64[k]) < TOL, k
            assert rel_err(grads[k], gold['g_' + k]) < TOL, k
    # refinable tilt angles / per-angle probe offset: full tables [3, n_theta] / [n_theta, 2], only this angle's entries are
    # non-zero.  Scalar sums with heavy cancellation: the reference's own fp32 values sit 1.4e-4 from fp64 (5e-4, see
    # test_oracle_golden.py)
    it = inp['cfg']['i_theta']
    if 'tilt_ls' in grads:
        assert rel_err(grads['tilt_ls'][:, it], grads64['tilt']) < 5e-4
        assert rel_err(grads['tilt_ls'], gold['g_tilt_ls']) < 5e-4
    if 'probe_pos_offset' in grads:
        assert rel_err(grads['probe_pos_offset'][it], grads64['probe_pos_offset']) < 5e-4
        assert rel_err(grads['probe_pos_offset'], gold['g_probe_pos_offset']) < 5e-4
    if 'slice_pos_cm_ls' in grads:
        assert rel_err(grads['slice_pos_cm_ls'], grads64['slice_pos_cm_ls']) < 5e-4
        assert rel_err(grads['slice_pos_cm_ls'], gold['g_slice_pos_cm_ls']) < 5e-4
    if 'prj_pos_offset' in grads:
        assert rel_err(grads['prj_pos_offset'][it], grads64['prj_pos_offset']) < 5e-4
        assert rel_err(grads['prj_pos_offset'], gold['g_prj_pos_offset']) < 5e-4


@pytest.mark.parametrize('P,sign,norm', [(256, 1, False), (512, 1, False), (128, -1, False), (128, -1, True)])
def test_large_probe_sizes_match_oracle(P, sign, norm):
    """Probe sizes 256 / 512 (2-D ptychography, two modes, probe + position refinement, far field) on the batched
    large-frame slice steps against the fp64 oracle (the 128-pixel cases also carry reference goldens); the sign -1
    convention (ifft2 detector transform, propagate.py:269-270) with and without normalize_fft."""
    CASES['_tmp_large'] = dict(CASES['ptycho2d_p128'], probe_size=(P, P), obj_size=(P + 40, P + 30, 1), seed=40 + P,
                               sign_convention=sign, normalize_fft=norm)
    try:
        inp, loss, pred, grads, fm = run_gpu_case('_tmp_large')
        _, loss64, pred64, grads64 = run_oracle_case('_tmp_large', torch.float64)
    finally:
        del CASES['_tmp_large']
    assert abs(loss - loss64) <= TOL * abs(loss64)
    assert rel_err(pred, pred64) < TOL
    for k in ('obj', 'probe_real', 'probe_imag', 'probe_pos_correction'):
        assert rel_err(grads[k], grads64[k]) < TOL, k


def test_single_batch_ptychography_model_matches_oracle():
    """SingleBatchPtychographyModel (forward_model.py:493-586: one probe position per call, no probe shift) on the fused
    kernels vs the oracle for single positions of the 3-D P = 72 case."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    name = 'ptycho3d_p72'
    inp, cfg, coord, reg = case_cfg(name)
    c = inp['cfg']
    n_theta = c['n_theta']
    P = c['probe_size']
    theta_ls = np.linspace(0, 2 * np.pi, n_theta, dtype='float32')
    theta_ls[c['i_theta']] = c['theta']
    prj_full = np.zeros([n_theta, 16, P[0], P[1]], dtype='float32')
    prj_full[c['i_theta'], inp['this_ind_batch']] = inp['prj']
    cv = _common_vars(inp, cfg, dev, prj_full, theta_ls, n_theta)
    cv['minibatch_size'] = 1
    fm = ab.SingleBatchPtychographyModel(loss_function_type=c['loss'], device=dev, common_vars_dict=cv,
                                         raw_data_type=c['raw_data_type'])
    t = lambda a, rg=True: torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=rg)
    for j in (0, 2):
        obj, pr, pi = t(inp['obj']), t(inp['probe_real']), t(inp['probe_imag'])
        args = dict(obj=obj, probe_real=pr, probe_imag=pi, probe_defocus_mm=None, probe_pos_offset=None,
                    this_i_theta=c['i_theta'], this_pos_batch=inp['pos'][j:j + 1], prj=prj_full,
                    probe_pos_correction=torch.zeros([n_theta, 16, 2], device=dev),
                    this_ind_batch=inp['this_ind_batch'][j:j + 1], tilt_ls=None, prj_pos_offset=None)
        diff = ab.Differentiator()
        diff.create_loss_node(fm.get_loss_function(), [0, fm.get_argument_index('probe_real')])
        g_obj, g_pr = diff.get_gradients(**args)
        l64, _, _, gr64 = orc.ptycho_loss_and_grads(torch.tensor(inp['obj'], dtype=torch.float64),
                                                    torch.tensor(inp['probe_real'], dtype=torch.float64),
                                                    torch.tensor(inp['probe_imag'], dtype=torch.float64), inp['pos'][j:j + 1],
                                                    torch.tensor(inp['prj'][j:j + 1], dtype=torch.float64), coord, cfg,
                                                    want=('obj', 'probe_real'))
        assert abs(fm.current_loss - float(l64)) <= TOL * abs(float(l64))
        assert rel_err(g_obj.cpu().numpy(), gr64['obj'].numpy()) < TOL
        assert rel_err(g_pr.cpu().numpy(), gr64['probe_real'].numpy()) < TOL


def test_per_pattern_loss_matches_oracle():
    """north_star: per-pattern loss within 1e-4 (the fused loss kernel reports one partial sum per pattern)."""
    name = 'ptycho3d_p72'
    inp, loss, pred, grads, fm = run_gpu_case(name)
    _, _, pred64, _ = run_oracle_case(name, torch.float64)
    lpp = fm.last_loss_per_pattern.cpu().numpy().astype('float64')
    ref = ((pred64 - inp['prj'].astype('float64')) ** 2).sum(axis=(1, 2))
    np.testing.assert_allclose(lpp, ref, rtol=TOL)


def test_multislice_propagate_batch_dropin():
    """The propagate.py:131-288 drop-in (explicit patch stack in, [real, imag] out) vs the oracle, incl. near field."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(3)
    for P, fp in ((64, 0), (72, 'inf'), (64, 'inf'), (64, 3e-4), (72, 3e-4), (72, None), (128, 'inf'), (256, 3e-4)):
        B, Nz = 3, 5
        grid = np.stack([rs.uniform(0, 2e-3, (B, P, P, Nz)), rs.uniform(0, 2e-4, (B, P, P, Nz))], -1).astype('float32')
        pr = rs.normal(size=(P, P)).astype('float32')
        pi = rs.normal(size=(P, P)).astype('float32')
        ref = orc.multislice_propagate_batch(torch.tensor(grid, dtype=torch.float64), torch.complex(
            torch.tensor(pr, dtype=torch.float64), torch.tensor(pi, dtype=torch.float64)), 5000., 1e-7, free_prop_cm=fp)
        re, im = ab.multislice_propagate_batch(torch.tensor(grid, device=dev), torch.tensor(pr, device=dev),
                                               torch.tensor(pi, device=dev), 5000., 1e-7, free_prop_cm=fp)
        out = torch.complex(re, im).cpu().numpy()
        assert rel_err(out, ref.numpy()) < TOL, (P, fp)


def test_near_field_gradients_match_oracle():
    """ADM_PROP_NEAR (finite free_prop_cm: propagate.py:271-280) forward + adjoint vs fp64 autograd."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(5)
    P, B, Nz = 64, 2, 4
    grid = np.stack([rs.uniform(0, 2e-3, (B, P, P, Nz)), rs.uniform(0, 2e-4, (B, P, P, Nz))], -1).astype('float32')
    pr = rs.normal(size=(P, P)).astype('float32')
    pi = rs.normal(size=(P, P)).astype('float32')
    wt = rs.normal(size=(B, P, P)).astype('float32')
    g64 = torch.tensor(grid, dtype=torch.float64, requires_grad=True)
    p64r = torch.tensor(pr, dtype=torch.float64, requires_grad=True)
    p64i = torch.tensor(pi, dtype=torch.float64, requires_grad=True)
    psi = orc.multislice_propagate_batch(g64, torch.complex(p64r, p64i), 5000., 1e-7, free_prop_cm=2e-4)
    l64 = ((psi.real ** 2 + psi.imag ** 2) * torch.tensor(wt, dtype=torch.float64)).sum()
    gg64, gr64, gi64 = torch.autograd.grad(l64, [g64, p64r, p64i])
    g = torch.tensor(grid, device=dev, requires_grad=True)
    r = torch.tensor(pr, device=dev, requires_grad=True)
    i = torch.tensor(pi, device=dev, requires_grad=True)
    re, im = ab.multislice_propagate_batch(g, r, i, 5000., 1e-7, free_prop_cm=2e-4)
    l = ((re ** 2 + im ** 2) * torch.tensor(wt, device=dev)).sum()
    gg, gr, gi = torch.autograd.grad(l, [g, r, i])
    assert abs(l.item() - l64.item()) <= TOL * abs(l64.item())
    assert rel_err(gg.cpu().numpy(), gg64.numpy()) < TOL
    assert rel_err(gr.cpu().numpy(), gr64.numpy()) < TOL
    assert rel_err(gi.cpu().numpy(), gi64.numpy()) < TOL


@pytest.mark.parametrize('P,fp', [(64, 'inf'), (64, 0), (72, 0), (72, 2e-4), (64, 2e-4)])
def test_three_pass_gradients_of_explicit_patch_stacks_match_oracle(P, fp):
    """Forward + adjoint of the three-pass kernels (ms3_kernels.cuh) through the propagate.py:131 drop-in for several
    wavefields: 64-pixel probes (64 = 4 x 16) with far-field, exit-wave and near-field detection, 72-pixel probes with
    exit-wave and near-field detection (the named cases cover 72 / far field and 64 / exit wave / one wavefield) vs fp64
    autograd."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(7)
    B, Nz = 3, 6
    grid = np.stack([rs.uniform(0, 2e-3, (B, P, P, Nz)), rs.uniform(0, 2e-4, (B, P, P, Nz))], -1).astype('float32')
    pr = rs.normal(size=(P, P)).astype('float32')
    pi = rs.normal(size=(P, P)).astype('float32')
    wt = rs.normal(size=(B, P, P)).astype('float32')
    g64 = torch.tensor(grid, dtype=torch.float64, requires_grad=True)
    p64r = torch.tensor(pr, dtype=torch.float64, requires_grad=True)
    p64i = torch.tensor(pi, dtype=torch.float64, requires_grad=True)
    psi = orc.multislice_propagate_batch(g64, torch.complex(p64r, p64i), 5000., 1e-7, free_prop_cm=fp)
    l64 = ((psi.real ** 2 + psi.imag ** 2) * torch.tensor(wt, dtype=torch.float64)).sum()
    gg64, gr64, gi64 = torch.autograd.grad(l64, [g64, p64r, p64i])
    g = torch.tensor(grid, device=dev, requires_grad=True)
    r = torch.tensor(pr, device=dev, requires_grad=True)
    i = torch.tensor(pi, device=dev, requires_grad=True)
    re, im = ab.multislice_propagate_batch(g, r, i, 5000., 1e-7, free_prop_cm=fp)
    l = ((re ** 2 + im ** 2) * torch.tensor(wt, device=dev)).sum()
    gg, gr, gi = torch.autograd.grad(l, [g, r, i])
    assert abs(l.item() - l64.item()) <= TOL * abs(l64.item())
    assert rel_err(gg.cpu().numpy(), gg64.numpy()) < TOL
    assert rel_err(gr.cpu().numpy(), gr64.numpy()) < TOL
    assert rel_err(gi.cpu().numpy(), gi64.numpy()) < TOL


def test_apply_rotation_and_transpose():
    """apply_rotation vs the oracle's grid_sample restatement, and <R x, y> == <x, R^T y>."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(7)
    Y, X, Z = 6, 48, 48
    obj = rs.normal(size=(Y, X, Z, 2)).astype('float32')
    for theta in (0.0, 0.3, 2.9, -1.1):
        lut = orc.rotation_coords([Y, X, Z], np.float32(theta))
        ref = orc.apply_rotation(torch.tensor(obj, dtype=torch.float64), lut).numpy()
        out = ab.apply_rotation(torch.tensor(obj, device=dev), lut)
        assert rel_err(out.cpu().numpy(), ref) < 1e-5, theta
        y = torch.tensor(rs.normal(size=(Y, X, Z, 2)).astype('float32'), device=dev)
        rty = ab.apply_rotation(y, lut, reverse=True)
        lhs = (out.double() * y.double()).sum().item()
        rhs = (torch.tensor(obj, device=dev).double() * rty.double()).sum().item()
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)


def test_rotation_lookup_matches_oracle_bitwise():
    from adorym_b200.util import rotation_lookup
    for theta in (0.0, 0.7, 3.1, 6.2):
        a = rotation_lookup([4, 80, 80], np.float32(theta))
        b = orc.rotation_coords([4, 80, 80], np.float32(theta))
        assert a.dtype == np.float16 and np.array_equal(a.view(np.uint16), b.view(np.uint16))


def test_adam_optimizer_matches_reference_golden():
    """AdamOptimizer.apply_gradient (fused kernel) vs the reference's three-call golden (repeated step index)."""
    import adorym_b200 as ab
    from make_golden import adam_inputs
    dev = torch.device('cuda', 0)
    gold = np.load(os.path.join(GOLD, 'adam.npz'))
    x, g_ls, steps = adam_inputs()
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': 1e-5})
    opt.create_container(list(x.shape), False, dev)
    xt = torch.tensor(x, device=dev)
    for i, (g, st) in enumerate(zip(g_ls, steps)):
        xt = opt.apply_gradient(xt, torch.tensor(g, device=dev), st, **opt.options_dict)
        np.testing.assert_allclose(xt.cpu().numpy(), gold['x%d' % i], rtol=2e-6, atol=1e-12)
        np.testing.assert_allclose(opt.params_whole_array_dict['m'].cpu().numpy(), gold['m%d' % i], rtol=2e-6, atol=1e-30)
        np.testing.assert_allclose(opt.params_whole_array_dict['v'].cpu().numpy(), gold['v%d' % i], rtol=2e-6, atol=1e-30)


def test_adam_epilogue_flags_and_mask():
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(9)
    x = torch.tensor(rs.normal(size=(5, 6, 7, 2)).astype('float32') * 1e-3, device=dev)
    g = torch.tensor(rs.normal(size=(5, 6, 7, 2)).astype('float32'), device=dev)
    mask = torch.tensor((rs.uniform(size=(5, 6, 7)) > 0.4).astype('float32'), device=dev)
    xr, _, _ = orc.adam_step(x.cpu(), g.cpu(), torch.zeros_like(x.cpu()), torch.zeros_like(x.cpu()), 0, step_size=1e-3)
    xr = xr.clamp(min=0)
    xr[..., 1] = 0
    xr = xr * mask.cpu().unsqueeze(-1)
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': 1e-3})
    out = opt.apply_gradient(x.clone(), g, 0, step_size=1e-3, clip_nonneg=True, zero_channel=1, mask=mask)
    np.testing.assert_allclose(out.cpu().numpy(), xr.numpy(), rtol=2e-6, atol=1e-12)


def test_gd_and_momentum_optimizers_match_reference_update():
    """GDOptimizer / MomentumOptimizer (fused adm_sgd_step) vs the reference's expressions evaluated by torch on the CPU
    (adorym/optimizers.py:432-464: x - step g with the halving schedule; :366-411: v = gamma v + step g, x = x - v),
    bit for bit, over three consecutive calls; plus the constraint epilogue."""
    import adorym_b200 as ab
    dev = torch.device('cuda', 0)
    rs = np.random.RandomState(4)
    x0 = torch.tensor(rs.normal(size=(6, 5, 7, 2)).astype('float32') * 1e-3)
    gs = [torch.tensor((rs.normal(size=(6, 5, 7, 2)) * 10.0 ** rs.uniform(-6, 0, size=(6, 5, 7, 2))).astype('float32'))
          for _ in range(3)]
    # gradient descent, dynamic rate: halved once past first_downrate_iteration
    opt = ab.GDOptimizer('obj', options_dict={'step_size': 1e-2})
    x, xr = x0.clone().to(dev), x0.clone()
    for i, g in zip((0, 5, 200), gs):
        x = opt.apply_gradient(x, g.to(dev), i, step_size=1e-2, dynamic_rate=True, first_downrate_iteration=92)
        ss = 1e-2 / 2. if i > 92 else 1e-2
        xr = xr - ss * g
        assert torch.equal(x.cpu(), xr)
    # momentum
    opt = ab.MomentumOptimizer('obj', options_dict={'step_size': 1e-2})
    opt.create_container(list(x0.shape), False, dev)
    x, xr, vr = x0.clone().to(dev), x0.clone(), torch.zeros_like(x0)
    for i, g in enumerate(gs):
        x = opt.apply_gradient(x, g.to(dev), i, step_size=1e-2, gamma=0.9)
        vr = 0.9 * vr + 1e-2 * g
        xr = xr - vr
        assert torch.equal(x.cpu(), xr)
        assert torch.equal(opt.params_whole_array_dict['v'].cpu(), vr)
    # epilogue: non-negativity, phase-only, mask
    mask = torch.tensor((rs.uniform(size=(6, 5, 7)) > 0.4).astype('float32'))
    x = opt.apply_gradient(x0.clone().to(dev), gs[0].to(dev), 3, step_size=1e-2, gamma=0.0, clip_nonneg=True, zero_channel=1,
                           mask=mask.to(dev))
    ref = (x0 - 1e-2 * gs[0]).clamp(min=0)
    ref[..., 1] = 0
    ref = ref * mask.unsqueeze(-1)
    assert torch.equal(x.cpu(), ref)


def test_full_size_properties_config_d():
    """BASELINE config D at full size (256^3, P=72, 529 positions, one angle): size-independent properties.
    (1) Parseval / unit-modulus transmission: with beta = 0, sum|det|^2 == P^2 sum|probe|^2 for every pattern.
    (2) Adjoint dot-product test of the whole chain w.r.t. the probe (linear in the probe): <A p, y> == <p, A^T y>.
    (3) Linearity of the gradient scatter: grad of 2 x batch == 2 x grad (same positions repeated)."""
    import adorym_b200 as ab
    from adorym_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(0)
    Y = X = Z = 256
    P = 72
    obj = torch.zeros((Y, X, Z, 2), device=dev)
    obj[..., 0] = torch.rand((Y, X, Z), device=dev, generator=g) * 1e-5
    pos = np.array([(y, x) for y in (np.arange(23) * 12) - 36 for x in (np.arange(23) * 12) - 36], dtype='int32')
    pos_t = torch.tensor(pos, device=dev)
    voxel_nm = np.array([1e-7] * 3) * 1e7
    h = ab.get_kernel(voxel_nm[-1], 1240. / 5000., voxel_nm, [P, P])
    plan = ops.MultislicePlan(P, Z, Y, X, ops.k1_of(5000., 1e-7), h, dev, free_prop='inf')
    lut = orc.rotation_coords([Y, X, Z], np.float32(0.4))
    rot = ops.RotateFn.apply(obj, ops.normalized_grid(lut, X, Z, dev))
    pr, pi = orc.gaussian_probe([P, P], 1, 6, 0.5, 6)
    probe = torch.tensor(np.stack([pr, pi], -1)[None].astype('float32'), device=dev, requires_grad=True)
    det = ops.MultisliceFn.apply(rot, None, probe, pos_t, plan, None, 'zyx')
    e_det = (det.double() ** 2).sum(dim=(1, 2, 3, 4))
    e_probe = (probe.double() ** 2).sum() * P * P
    assert torch.allclose(e_det, e_probe.expand_as(e_det), rtol=1e-4)
    y = torch.randn(det.shape, device=dev, generator=g)
    (gp,) = torch.autograd.grad((det * y).sum(), [probe])
    lhs = (det.double() * y.double()).sum().item()
    rhs = (probe.double() * gp.double()).sum().item()
    # <A p, y> is a random-sign sum over 529 * 72 * 72 * 2 terms: the tolerance is relative to the natural scale of such
    # a sum (its root-sum-square), not to the heavily cancelled total
    scale = (det.double() * y.double()).norm().item()
    assert abs(lhs - rhs) <= 1e-4 * scale
    # gradient scatter linearity on a sub-batch
    rot2 = rot.detach().requires_grad_(True)
    sub = pos_t[100:140].contiguous()
    d1 = ops.MultisliceFn.apply(rot2, None, probe.detach(), sub, plan, None, 'zyx')
    (g1,) = torch.autograd.grad((d1 * y[100:140]).sum(), [rot2])
    sub2 = torch.cat([sub, sub]).contiguous()
    d2 = ops.MultisliceFn.apply(rot2, None, probe.detach(), sub2, plan, None, 'zyx')
    (g2,) = torch.autograd.grad((d2 * torch.cat([y[100:140], y[100:140]])).sum(), [rot2])
    assert rel_err(g2.cpu().numpy(), 2 * g1.cpu().numpy()) < 1e-5


def _config_d_model(dev, pos, prj, theta, reg=True):
    """PtychographyModel of BASELINE config D at full size (256^3, P=72, 256 slices, far field, lsq + L1 + TV)."""
    import adorym_b200 as ab
    Y = X = Z = 256
    P = [72, 72]
    voxel_nm = np.array([1e-7] * 3) * 1e7
    h = ab.get_kernel(voxel_nm[-1], 1240. / 5000., voxel_nm, P)
    theta_ls = np.array([theta], dtype='float32')
    cv = dict(unknown_type='delta_beta', normalize_fft=False, sign_convention=1, rotate_out_of_loop=False,
              scale_ri_by_k=True, is_minus_logged=False, forward_algorithm='fresnel', stdout_options={},
              poisson_multiplier=1., common_probe_pos=True, binning=1, prj=prj, device_obj=dev, probe_size=P,
              fresnel_approx=True, two_d_mode=False, minibatch_size=len(pos), ds_level=1, this_obj_size=[Y, X, Z],
              energy_ev=5000., psize_cm=1e-7, h=h, pure_projection=False, n_dp_batch=20, free_prop_cm='inf',
              optimize_probe_defocusing=False, optimize_probe_pos_offset=False, optimize_prj_pos_offset=False,
              optimize_all_probe_pos=False, optimize_tilt=False, initial_tilt=None, n_probe_modes=1, n_theta=1,
              precalculate_rotation_coords=False, theta_ls=theta_ls, theta_downsample=None, beamstop=None)
    fm = ab.PtychographyModel(loss_function_type='lsq', device=dev, common_vars_dict=cv, raw_data_type='magnitude')
    if reg:
        fm.add_regularizers([ab.L1Regularizer(1e-9 * 1.7e7, 1e-10 * 1.7e7), ab.TVRegularizer(1e-9 * 1.7e7)])
    return fm, h


def test_config_d_full_size_matches_oracle(monkeypatch):
    """BASELINE config D at FULL size (256^3 object, P = 72, 256 slices, far field, lsq + L1 + TV, rotation).
    (a) six of the 529 positions (two overhanging corners, the centre, interior ones): loss, per-pattern loss and
        magnitudes vs the oracle evaluated in fp64 on the same inputs, 1e-4 relative.  Object gradient: through 256
        slices dL/ddelta = -k1 Im(g conj(psi)) cancels heavily and the reference algorithm's OWN fp32 evaluation sits
        4.9e-3 (delta channel 1.1e-2) away from its fp64 evaluation (measured with the oracle at this size), so the bar
        is: at least as close to the fp64 gradient as 1.2x the reference algorithm in fp32 (the kernels carry T - 1
        instead of T and land at 3.8e-3);
    (b) all 529 positions of the angle (the benchmark's minibatch, which runs on the persistent balanced schedule with
        wavefields migrating between CTAs): per-pattern loss and magnitudes of the six positions vs the oracle, and the
        whole gradient vs the one-CTA-per-wavefield launch of the same kernels (ADM_MS_GRID forces it)."""
    import adorym_b200 as ab
    from cases import _smooth_noise
    dev = torch.device('cuda', 0)
    Y = X = Z = 256
    P = 72
    rs = np.random.RandomState(5)
    obj = np.stack([_smooth_noise(rs, (Y, X, Z), 1e-5), _smooth_noise(rs, (Y, X, Z), 1e-6)], -1)
    pos_all = np.array([(y, x) for y in (np.arange(23) * 12) - 36 for x in (np.arange(23) * 12) - 36], dtype='int64')
    sel = np.array([0, 17, 22, 264, 300, 528])
    theta = np.float32(0.83)
    pr, pi = orc.gaussian_probe([P, P], 1, 6, 0.5, 6)
    pr, pi = pr[None].astype('float32'), pi[None].astype('float32')
    amp = float(np.sqrt((pr ** 2 + pi ** 2).sum())) * 2.0
    prj_all = (amp * rs.uniform(0.2, 1.0, size=(529, P, P))).astype('float32')
    reg = dict(alpha_d=1e-9 * 1.7e7, alpha_b=1e-10 * 1.7e7, gamma=1e-9 * 1.7e7)

    def gpu_run(ind):
        fm, h = _config_d_model(dev, pos_all[ind], prj_all[None], theta)
        o = torch.tensor(obj, device=dev, requires_grad=True)
        args = dict(obj=o, probe_real=torch.tensor(pr, device=dev), probe_imag=torch.tensor(pi, device=dev),
                    probe_defocus_mm=None, probe_pos_offset=None, this_i_theta=0, this_pos_batch=pos_all[ind],
                    prj=prj_all[None], probe_pos_correction=torch.zeros([1, 529, 2], device=dev), this_ind_batch=ind,
                    tilt_ls=None, prj_pos_offset=None)
        diff = ab.Differentiator()
        diff.create_loss_node(fm.get_loss_function(), [0])
        (g,) = diff.get_gradients(**args)
        torch.cuda.synchronize()
        return fm.current_loss, fm.last_pred.cpu().numpy(), fm.last_loss_per_pattern.cpu().numpy(), g, h

    # ---- (a) six positions vs the fp64 oracle ----
    loss6, pred6, lpp6, g6, h = gpu_run(sel)
    cfg = dict(this_obj_size=[Y, X, Z], probe_size=[P, P], energy_ev=5000., psize_cm=1e-7, h=h, free_prop_cm='inf',
               unknown_type='delta_beta')
    lut = orc.rotation_coords([Y, X, Z], theta)
    l64, _, p64, gr64 = orc.ptycho_loss_and_grads(torch.tensor(obj, dtype=torch.float64), torch.tensor(pr, dtype=torch.float64),
                                                  torch.tensor(pi, dtype=torch.float64), pos_all[sel],
                                                  torch.tensor(prj_all[sel], dtype=torch.float64), lut, cfg, reg=reg,
                                                  want=('obj',))
    p64 = p64.numpy()
    assert abs(loss6 - float(l64)) <= TOL * abs(float(l64))
    ref_lpp = ((p64 - prj_all[sel].astype('float64')) ** 2).sum(axis=(1, 2))
    np.testing.assert_allclose(lpp6, ref_lpp, rtol=TOL)
    for b in range(len(sel)):
        assert rel_err(pred6[b], p64[b]) < TOL, 'pattern %d' % sel[b]
    _, _, _, gr32 = orc.ptycho_loss_and_grads(torch.tensor(obj), torch.tensor(pr), torch.tensor(pi), pos_all[sel],
                                              torch.tensor(prj_all[sel]), lut, cfg, reg=reg, want=('obj',))
    ref32_err = rel_err(gr32['obj'].numpy(), gr64['obj'].numpy())
    gpu_err = rel_err(g6.cpu().numpy(), gr64['obj'].numpy())
    print('config D full size: |g_gpu - g_64| / |g_64| = %.3e, reference algorithm in fp32: %.3e' % (gpu_err, ref32_err))
    assert gpu_err < max(TOL, 1.2 * ref32_err)
    del g6, gr32

    # ---- (b) the whole angle on the balanced schedule ----
    ind = np.arange(529)
    loss_b, pred_b, lpp_b, g_b, _ = gpu_run(ind)
    np.testing.assert_allclose(lpp_b[sel], ref_lpp, rtol=TOL)
    for j, b in enumerate(sel):
        assert rel_err(pred_b[b], p64[j]) < TOL, 'pattern %d' % b
    monkeypatch.setenv('ADM_MS_GRID', '100000')          # every wavefield gets its own CTA (no hand-off)
    loss_c, pred_c, lpp_c, g_c, _ = gpu_run(ind)
    monkeypatch.delenv('ADM_MS_GRID')
    assert np.array_equal(pred_b, pred_c)                 # the forward arithmetic per wavefield is identical
    assert abs(loss_b - loss_c) <= 1e-6 * abs(loss_c)
    d = (g_b - g_c).double().norm().item() / g_c.double().norm().item()
    assert d < 1e-5, d                                    # (order of the L2 reduce-adds differs)


@pytest.mark.parametrize('name,grid', [('ptycho3d_p72_edge', 3), ('ptycho3d_p72_edge', 7), ('ptycho3d_p72', 3),
                                       ('ptycho2d_modes', 5), ('ptycho2d_realimag_p64', 4)])
def test_balanced_schedule_with_handoff_matches_oracle(name, grid, monkeypatch):
    """Wavefields cut at arbitrary slices and handed from CTA to CTA (ADM_MS_GRID forces fewer co-resident CTAs than
    wavefields): same parity bar as the default launch."""
    monkeypatch.setenv('ADM_MS_GRID', str(grid))
    inp, loss, pred, grads, fm = run_gpu_case(name)
    _, loss64, pred64, grads64 = run_oracle_case(name, torch.float64)
    assert abs(loss - loss64) <= TOL * abs(loss64)
    assert rel_err(pred, pred64) < TOL
    for k in grads:
        assert rel_err(grads[k], grads64[k]) < TOL, k


@pytest.mark.parametrize('name', ['ptycho3d_p72', 'ptycho3d_p72_edge', 'ptycho3d_p72_exitwave', 'sparse3_p72'])
def test_five_pass_kernels_for_72_pixel_probes_match_oracle(name, monkeypatch):
    """P = 72 with a separable transfer function runs on the three-pass kernels (ms3_kernels.cuh) by default; the five-pass
    kernels of multislice.cu remain the path for near-field detection, general H and the adjoint-wavefield checkpoints,
    so they keep the same parity bar (ADM_MS3=0 selects them for every shape)."""
    monkeypatch.setenv('ADM_MS3', '0')
    inp, loss, pred, grads, fm = run_gpu_case(name)
    _, loss64, pred64, grads64 = run_oracle_case(name, torch.float64)
    assert abs(loss - loss64) <= TOL * abs(loss64)
    assert rel_err(pred, pred64) < TOL
    assert rel_err(grads['obj'], grads64['obj']) < TOL
    for k in ('probe_real', 'probe_imag'):
        if k in grads:
            assert rel_err(grads[k], grads64[k]) < TOL, k


def test_reconstruct_ptychography_matches_oracle_driver():
    """End-to-end: three minibatch updates of reconstruct_ptychography (3-D, rotation, far field, Adam, L1+TV) vs the
    same schedule evaluated with the oracle.  Tolerance for a reconstruction after updates (north_star: "within a stated
    tolerance"): the first loss (no update yet) within 1e-4; every later loss within 1e-4 + five times the distance between the
    reference algorithm's OWN fp32 and fp64 evaluations of that loss -- Adam's first steps are sign-like
    (m / (sqrt(v) + eps) ~ sign(g) where |g| >> eps), so roundoff in near-zero gradient voxels flips whole steps and the
    fp32 reference itself sits ~2e-3 away from the fp64 one after a single update at this size."""
    import adorym_b200 as ab
    from oracle_driver import oracle_reconstruct
    rs = np.random.RandomState(21)
    Y = X = Z = 80
    P = 72
    n_theta = 3
    pos = [(y, x) for y in (-20, 4) for x in (-12, 8)]
    from cases import _smooth_noise
    truth = np.stack([_smooth_noise(rs, (Y, X, Z), 6e-4), _smooth_noise(rs, (Y, X, Z), 6e-5)], -1)
    init = [_smooth_noise(rs, (Y, X, Z), 3e-4), _smooth_noise(rs, (Y, X, Z), 3e-5)]
    params = dict(obj_size=(Y, X, Z), probe_pos=pos, theta_st=0, theta_end=2 * np.pi, n_theta=n_theta, energy_ev=5000.,
                  psize_cm=1e-7, free_prop_cm='inf', minibatch_size=2, n_epochs=1, alpha_d=1e-9 * 1.7e7,
                  alpha_b=1e-10 * 1.7e7, gamma=1e-9 * 1.7e7, learning_rate=5e-5, probe_type='gaussian',
                  probe_mag_sigma=6, probe_phase_sigma=6, probe_phase_max=0.5, initial_guess=init)
    data, hist_ref, obj_ref = oracle_reconstruct(truth, max_batches=3, **params)
    _, hist_64, _ = oracle_reconstruct(truth, max_batches=3, dtype=torch.float64, **params)
    out = ab.reconstruct_ptychography(data, **params, t_max_min=None, verbose=False, save_outputs=False,
                                      output_folder='/tmp/adm_test_recon', max_batches_for_test=3)
    hist = out['loss_history']
    assert len(hist) >= 3
    assert abs(hist[0][2] - hist_ref[0]) <= 1e-4 * abs(hist_ref[0])
    for (e, b, l, _), l32, l64 in zip(hist[:3], hist_ref[:3], hist_64[:3]):
        # (a second fp32 evaluation of the same algorithm -- different summation orders -- moves by several times the
        # fp32-vs-fp64 distance of the first: Adam's sign-like first steps, see tests/test_driver_golden.py)
        assert abs(l - l64) <= 1e-4 * abs(l64) + 5 * abs(l32 - l64), (b, l, l32, l64)
