"""Golden vectors for the MultiDistModel cases (cases.MD_CASES) from the UNMODIFIED reference (PyTorch backend, CPU):

    PYTHONPATH=oracle/ref_shims:/root/reference python tests/golden/make_golden_multidist.py

Stored per case: loss, predicted magnitudes (every 4th pixel), the object gradient compacted as in make_golden.py,
the probe gradient (every 4th pixel) and the per-distance shift gradient."""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import MD_CASES, make_md_inputs, md_tile_positions  # noqa: E402
from make_golden import compact_grad  # noqa: E402


def run_reference(name):
    import torch
    import adorym  # noqa: F401
    import adorym.global_settings as gs
    import adorym.wrappers as w
    from adorym.forward_model import MultiDistModel
    gs.backend = 'pytorch'
    inp = make_md_inputs(name)
    c = inp['cfg']
    N, M, D = c['N'], c['n_modes'], len(c['dists'])
    voxel_nm = np.array([c['psize_cm']] * 3) * 1e7
    lmbda_nm = 1240. / c['energy_ev']
    prj_full = inp['prj'][None]                  # [n_theta = 1, n_dists, N, N]
    tiles = c.get('tiles')
    sub = tiles['sub'] if tiles else N
    szw = tiles['szw'] if tiles else 0
    ind = np.array(tiles['ind']) if tiles else np.array([0])
    pos = md_tile_positions(N, sub)[ind] if tiles else np.array([[0, 0]])
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        cv = dict(unknown_type=c['unknown_type'], normalize_fft=False, sign_convention=1, rotate_out_of_loop=False,
                  scale_ri_by_k=True, is_minus_logged=False, forward_algorithm='fresnel',
                  stdout_options={'save_stdout': False, 'output_folder': tmp, 'timestamp': ''}, poisson_multiplier=1.,
                  common_probe_pos=True, binning=1, prj=prj_full, device_obj=None, lmbda_nm=lmbda_nm, voxel_nm=voxel_nm,
                  probe_size=[N, N], subprobe_size=[sub, sub], fresnel_approx=True, two_d_mode=True, minibatch_size=len(ind),
                  ds_level=1, this_obj_size=[N, N, 1], energy_ev=c['energy_ev'], psize_cm=c['psize_cm'], h=None,
                  pure_projection=False, n_dp_batch=tiles['n_dp_batch'] if tiles else 1, free_prop_cm=c['dists'], optimize_probe_defocusing=False,
                  optimize_probe_pos_offset=False, optimize_prj_pos_offset=False, optimize_all_probe_pos=c['shift'],
                  optimize_free_prop=c.get('free_prop', False), optimize_prj_affine=c.get('affine', False), optimize_ctf_lg_kappa=False, optimize_tilt=False,
                  initial_tilt=None, debug=False, output_folder=tmp, n_probe_modes=M, n_theta=1,
                  precalculate_rotation_coords=False, theta_ls=np.zeros(1, 'float32'), theta_downsample=None,
                  beamstop=None, u_free=None, v_free=None)
        if c.get('free_prop', False):
            from adorym.propagate import gen_freq_mesh
            uf, vf = gen_freq_mesh(voxel_nm, [N, N])              # ptychography.py:320-323
            cv['u_free'], cv['v_free'] = w.create_variable(uf, requires_grad=False), w.create_variable(vf, requires_grad=False)
        fm = MultiDistModel(loss_function_type=c['loss'], distribution_mode=None, device=None, common_vars_dict=cv,
                            raw_data_type=c['raw_data_type'])
        obj = w.create_variable(inp['obj'], requires_grad=True)
        pr = w.create_variable(inp['probe_real'], requires_grad=True)
        pi = w.create_variable(inp['probe_imag'], requires_grad=True)
        ppc = w.create_variable(inp['probe_pos_correction'] if c['shift'] else np.zeros([D, 2]), requires_grad=c['shift'])
        args = dict(obj=obj, probe_real=pr, probe_imag=pi, probe_defocus_mm=w.create_variable(0.0),
                    probe_pos_offset=w.zeros([1, 2], requires_grad=True), this_i_theta=0,
                    this_pos_batch=pos, prj=prj_full, probe_pos_correction=ppc, this_ind_batch=ind,
                    free_prop_cm=c['dists'], safe_zone_width=szw, prj_affine_ls=None, ctf_lg_kappa=None,
                    prj_pos_offset=w.zeros([1, 2], requires_grad=True))
        if c.get('free_prop', False):
            args['free_prop_cm'] = w.create_variable(np.array(c['dists']), requires_grad=True)
        if c.get('affine', False):
            args['prj_affine_ls'] = w.create_variable(inp['prj_affine_ls'], requires_grad=True)
        opt_args, names = [0], ['obj']
        if c.get('free_prop', False):
            opt_args += [fm.get_argument_index('free_prop_cm')]
            names += ['free_prop_cm']
        if c['optimize_probe']:
            opt_args += [fm.get_argument_index('probe_real'), fm.get_argument_index('probe_imag')]
            names += ['probe_real', 'probe_imag']
        if c['shift']:
            opt_args += [fm.get_argument_index('probe_pos_correction')]
            names += ['probe_pos_correction']
        if c.get('affine', False):
            opt_args += [fm.get_argument_index('prj_affine_ls')]
            names += ['prj_affine_ls']
        grads = w.get_gradients(fm.get_loss_function(), opt_args_ls=opt_args, **args)
        loss = float(fm.current_loss)
        with torch.no_grad():
            pred = fm.predict(**args)
        if tiles:
            # the reference's single-mode magnitude is torch.sqrt(ex_int) (forward_model.py:1017): not finite-gradient at 0
            assert np.isfinite(loss)
        out = {'loss': np.float64(loss), 'pred_sub': pred.numpy()[:, ::4, ::4].astype('float32'),
               'pred_norm': np.float64(np.sqrt((pred.numpy().astype('float64') ** 2).sum()))}
        for n, g in zip(names, grads):
            g = g.detach().numpy()
            if n == 'obj':
                for k, v in compact_grad(g).items():
                    out['gobj_' + k] = v
            elif n in ('probe_pos_correction', 'free_prop_cm', 'prj_affine_ls'):
                out['g_' + n] = g.astype('float32')
            else:
                out['g_' + n + '_sub'] = g[:, ::4, ::4].astype('float32')
                out['g_' + n + '_norm'] = np.float64(np.sqrt((g.astype('float64') ** 2).sum()))
        return out
    finally:
        os.chdir(cwd)


if __name__ == '__main__':
    for name in (sys.argv[1:] or list(MD_CASES.keys())):
        out = run_reference(name)
        fn = os.path.join(HERE, name + '.npz')
        np.savez_compressed(fn, **out)
        print(name, '-> loss %.8e' % out['loss'], os.path.getsize(fn) // 1024, 'KB')
