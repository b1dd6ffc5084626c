"""Generate golden vectors by running the UNMODIFIED reference (mdw771/adorym, PyTorch backend on CPU)
on the seeded inputs of cases.py.  Run in the build container only:

    PYTHONPATH=oracle/ref_shims:/root/reference python tests/golden/make_golden.py

The reference is imported from /root/reference (never copied); the I/O shims only make
`import adorym` possible (h5py / dxchange / matplotlib are not installed here).
What is stored per case (compact, a few 100 KB):  loss, predicted magnitudes, probe / position
gradients in full, and for the object gradient its three axis projections, 4096 sampled voxels
and its L2 norm (the full inputs are regenerated from the seed by cases.make_inputs).
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import CASES, make_inputs  # noqa: E402


def compact_grad(g, seed=0):
    g = np.asarray(g, dtype='float64')
    rs = np.random.RandomState(seed)
    flat = g.reshape(-1)
    idx = rs.randint(0, flat.size, size=4096)
    out = {'norm': np.sqrt((flat ** 2).sum()), 'sample_idx': idx.astype('int64'), 'sample': flat[idx].astype('float32')}
    for ax in range(3):
        out['proj%d' % ax] = g.sum(axis=ax).astype('float32')
    return out


def run_reference(name):
    import torch
    import adorym
    import adorym.global_settings as gs
    import adorym.wrappers as w
    from adorym.forward_model import PtychographyModel, SingleBatchFullfieldModel, SparseMultisliceModel
    from adorym.propagate import gen_freq_mesh
    from adorym.propagate import get_kernel
    from adorym.util import save_rotation_lookup
    from adorym.regularizers import L1Regularizer, TVRegularizer, ReweightedL1Regularizer
    gs.backend = 'pytorch'
    inp = make_inputs(name)
    c = inp['cfg']
    Y, X, Z = c['obj_size']
    P = list(c['probe_size'])
    n_theta = c['n_theta']
    two_d = (Z == 1)
    theta_ls = np.linspace(0, 2 * np.pi, n_theta, dtype='float32')
    theta_ls[c['i_theta']] = c['theta']
    sign = c.get('sign_convention', 1)
    voxel_nm = np.array([c['psize_cm']] * 3) * 1e7
    lmbda_nm = 1240. / c['energy_ev']
    h = get_kernel(voxel_nm[-1] * c.get('binning', 1), lmbda_nm, voxel_nm, P, fresnel_approx=True, sign_convention=sign)
    B = len(inp['pos'])
    prj_full = np.zeros([n_theta, 16, P[0], P[1]], dtype='float32')
    prj_full[c['i_theta'], inp['this_ind_batch']] = inp['prj']
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        # the reference reads the LUT file even in two_d_mode (forward_model.py:233-235)
        os.makedirs('arrsize_{}_{}_{}_ntheta_{}'.format(Y, X, Z, n_theta))
        save_rotation_lookup([Y, X, Z], theta_ls)
        cv = dict(unknown_type=c['unknown_type'], normalize_fft=c.get('normalize_fft', False), sign_convention=sign,
                  rotate_out_of_loop=c.get('rotate_out_of_loop', False), scale_ri_by_k=True, is_minus_logged=False, forward_algorithm='fresnel',
                  stdout_options={'save_stdout': False, 'output_folder': tmp, 'timestamp': ''},
                  poisson_multiplier=c.get('poisson_multiplier', 1.), common_probe_pos=True, binning=c.get('binning', 1),
                  prj=prj_full,
                  device_obj=None, lmbda_nm=lmbda_nm, voxel_nm=voxel_nm, probe_size=P, fresnel_approx=True,
                  two_d_mode=two_d, minibatch_size=B, ds_level=1, this_obj_size=[Y, X, Z], energy_ev=c['energy_ev'],
                  psize_cm=c['psize_cm'], h=h, pure_projection=c.get('pure_projection', False), n_dp_batch=20, free_prop_cm=c['free_prop_cm'],
                  optimize_probe_defocusing=False, optimize_probe_pos_offset=c.get('optimize_probe_pos_offset', False),
                  optimize_prj_pos_offset=c.get('optimize_prj_pos_offset', False),
                  optimize_all_probe_pos=c.get('optimize_all_probe_pos', False), optimize_tilt=c.get('optimize_tilt', False),
                  initial_tilt=None,
                  debug=False, output_folder=tmp, n_probe_modes=c['n_modes'], n_theta=n_theta,
                  precalculate_rotation_coords=True, theta_ls=theta_ls, theta_downsample=None,
                  beamstop=None if inp['beamstop'] is None else torch.tensor(inp['beamstop'].copy()))
        Model = {'fullfield': SingleBatchFullfieldModel, 'sparse': SparseMultisliceModel}.get(c['kind'], PtychographyModel)
        if c['kind'] == 'sparse':
            u, v = gen_freq_mesh(voxel_nm, P)                    # ptychography.py:289-292
            cv['u'], cv['v'] = w.create_variable(u, requires_grad=False), w.create_variable(v, requires_grad=False)
        fm = Model(loss_function_type=c['loss'], distribution_mode=None, device=None, common_vars_dict=cv,
                   raw_data_type=c['raw_data_type'])
        regs = []
        if c.get('alpha_d') and c.get('reweighted', False):
            r = ReweightedL1Regularizer(c['alpha_d'], c['alpha_b'], unknown_type=c['unknown_type'])
            r.update_l1_weight(torch.tensor(inp['weight_l1']))
            regs.append(r)
        elif c.get('alpha_d'):
            regs.append(L1Regularizer(c['alpha_d'], c['alpha_b'], unknown_type=c['unknown_type']))
        if c.get('gamma'):
            regs.append(TVRegularizer(c['gamma'], unknown_type=c['unknown_type']))
        fm.add_regularizers(regs)
        obj = w.create_variable(inp['obj'], requires_grad=True)
        pr = w.create_variable(inp['probe_real'], requires_grad=True)
        pi = w.create_variable(inp['probe_imag'], requires_grad=True)
        if inp['probe_pos_correction'] is not None:
            ppc = w.create_variable(inp['probe_pos_correction'], requires_grad=True)
        else:
            ppc = w.create_variable(np.zeros([n_theta, 16, 2]), requires_grad=False)
        args = dict(obj=obj, probe_real=pr, probe_imag=pi, probe_defocus_mm=w.create_variable(0.0),
                    probe_pos_offset=w.zeros([n_theta, 2], requires_grad=True), this_i_theta=c['i_theta'],
                    this_pos_batch=inp['pos'], prj=prj_full, probe_pos_correction=ppc,
                    this_ind_batch=inp['this_ind_batch'], tilt_ls=np.zeros([3, n_theta]),
                    prj_pos_offset=w.zeros([n_theta, 2], requires_grad=True))
        if c.get('optimize_tilt', False):
            tilt_full = np.zeros([3, n_theta])
            tilt_full[0] = theta_ls                                   # ptychography.py:680-682
            tilt_full[:, c['i_theta']] = c['tilt']
            args['tilt_ls'] = w.create_variable(tilt_full, requires_grad=True)
        if c.get('optimize_probe_pos_offset', False):
            off = np.zeros([n_theta, 2])
            off[c['i_theta']] = c['probe_pos_offset']
            args['probe_pos_offset'] = w.create_variable(off, requires_grad=True)
        if c.get('optimize_prj_pos_offset', False):
            off = np.zeros([n_theta, 2])
            off[c['i_theta']] = c['prj_pos_offset']
            args['prj_pos_offset'] = w.create_variable(off, requires_grad=True)
        if c['kind'] == 'sparse':
            del args['tilt_ls']
            args['slice_pos_cm_ls'] = w.create_variable(np.array(c['slice_pos_cm']), requires_grad=c.get('optimize_slice_pos', False))
        args = {k: args[k] for k in fm.argument_ls}            # w.get_gradients indexes the kwargs BY POSITION
        opt_args = [0]
        names = ['obj']
        if c.get('optimize_probe', False):
            opt_args += [fm.get_argument_index('probe_real'), fm.get_argument_index('probe_imag')]
            names += ['probe_real', 'probe_imag']
        if c.get('optimize_probe_pos_offset', False):
            opt_args += [fm.get_argument_index('probe_pos_offset')]
            names += ['probe_pos_offset']
        if c.get('optimize_prj_pos_offset', False):
            opt_args += [fm.get_argument_index('prj_pos_offset')]
            names += ['prj_pos_offset']
        if c.get('optimize_all_probe_pos', False):
            opt_args += [fm.get_argument_index('probe_pos_correction')]
            names += ['probe_pos_correction']
        if c.get('optimize_slice_pos', False):
            opt_args += [fm.get_argument_index('slice_pos_cm_ls')]
            names += ['slice_pos_cm_ls']
        if c.get('optimize_tilt', False):
            opt_args += [fm.get_argument_index('tilt_ls')]
            names += ['tilt_ls']
        loss_fn = fm.get_loss_function()
        grads = w.get_gradients(loss_fn, opt_args_ls=opt_args, **args)
        loss = float(fm.current_loss)
        with torch.no_grad():
            pred = fm.predict(**args)
            if isinstance(pred, (tuple, list)):
                pred = pred[0]
        out = {'loss': np.float64(loss), 'pred': pred.numpy().astype('float32')}
        for n, g in zip(names, grads):
            g = g.detach().numpy()
            if n == 'obj':
                for k, v in compact_grad(g).items():
                    out['gobj_' + k] = v
            else:
                out['g_' + n] = g.astype('float32')
        return out
    finally:
        os.chdir(cwd)


def adam_inputs():
    rs = np.random.RandomState(21)
    shape = (12, 10, 14, 2)
    x = rs.normal(size=shape).astype('float32') * 1e-3
    gs_ = [(rs.normal(size=shape) * 10.0 ** rs.uniform(-9, 1, size=shape)).astype('float32') for _ in range(3)]
    return x, gs_, [0, 0, 1]   # step indices: the reference repeats an index within an angle (quirk 4)


def run_adam():
    """AdamOptimizer.apply_gradient (adorym/optimizers.py:275-337) for three consecutive calls."""
    import torch
    import adorym
    import adorym.global_settings as gs
    gs.backend = 'pytorch'
    x, g_ls, steps = adam_inputs()
    opt = adorym.AdamOptimizer('obj', output_folder='.', options_dict={'step_size': 1e-5})
    opt.create_container(list(x.shape), False, None)
    xt = torch.tensor(x)
    out = {}
    with torch.no_grad():
        for i, (g, st) in enumerate(zip(g_ls, steps)):
            xt = opt.apply_gradient(xt, torch.tensor(g), st, step_size=1e-5)
            out['x%d' % i] = xt.numpy().copy()
            out['m%d' % i] = opt.params_whole_array_dict['m'].numpy().copy()
            out['v%d' % i] = opt.params_whole_array_dict['v'].numpy().copy()
    return out


def main():
    names = sys.argv[1:] or (list(CASES.keys()) + ['adam'])
    if 'adam' in names:
        names.remove('adam')
        out = run_adam()
        np.savez_compressed(os.path.join(HERE, 'adam.npz'), **out)
        print('adam ->', os.path.getsize(os.path.join(HERE, 'adam.npz')) // 1024, 'KB')
    for name in names:
        out = run_reference(name)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'loss', out['loss'], '->', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
