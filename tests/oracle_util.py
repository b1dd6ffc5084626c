"""Helpers shared by the oracle tests and the GPU parity tests (test infrastructure)."""
import numpy as np
import torch
from cases import make_inputs
from oracle import multislice_oracle as orc


def compact_grad(g, seed=0):
    g = np.asarray(g, dtype='float64')
    rs = np.random.RandomState(seed)
    flat = g.reshape(-1)
    idx = rs.randint(0, flat.size, size=4096)
    out = {'norm': np.sqrt((flat ** 2).sum()), 'sample_idx': idx.astype('int64'), 'sample': flat[idx].astype('float32')}
    for ax in range(3):
        out['proj%d' % ax] = g.sum(axis=ax).astype('float32')
    return out


def rel_err(a, b):
    dt = 'complex128' if (np.iscomplexobj(a) or np.iscomplexobj(b)) else 'float64'
    a = np.asarray(a, dtype=dt)
    b = np.asarray(b, dtype=dt)
    return float(np.sqrt((np.abs(a - b) ** 2).sum()) / (np.sqrt((np.abs(b) ** 2).sum()) + 1e-300))


def case_cfg(name):
    inp = make_inputs(name)
    c = inp['cfg']
    Y, X, Z = c['obj_size']
    sign = c.get('sign_convention', 1)
    voxel_nm = np.array([c['psize_cm']] * 3) * 1e7
    lmbda_nm = 1240. / c['energy_ev']
    h = orc.get_kernel(voxel_nm[-1] * c.get('binning', 1), lmbda_nm, voxel_nm, c['probe_size'], fresnel_approx=True,
                       sign_convention=sign)
    cfg = dict(this_obj_size=[Y, X, Z], probe_size=list(c['probe_size']), energy_ev=c['energy_ev'], psize_cm=c['psize_cm'],
               h=h, free_prop_cm=c['free_prop_cm'], unknown_type=c['unknown_type'], binning=c.get('binning', 1),
               pure_projection=c.get('pure_projection', False),
               normalize_fft=c.get('normalize_fft', False), sign_convention=sign, scale_ri_by_k=True,
               fresnel_approx=True, optimize_all_probe_pos=c.get('optimize_all_probe_pos', False), two_d_mode=(Z == 1),
               loss_function_type=c['loss'], raw_data_type=c['raw_data_type'],
               poisson_multiplier=c.get('poisson_multiplier', 1.))
    coord = None if Z == 1 else orc.rotation_coords([Y, X, Z], np.float32(c['theta']))
    reg = None
    if c.get('alpha_d') or c.get('gamma'):
        reg = dict(alpha_d=c.get('alpha_d'), alpha_b=c.get('alpha_b'), gamma=c.get('gamma'))
        if inp.get('weight_l1') is not None:
            reg['weight_l1'] = inp['weight_l1']
    cfg['beamstop'] = inp.get('beamstop')
    if c.get('slice_pos_cm') is not None:
        cfg['slice_pos_cm_ls'] = np.array(c['slice_pos_cm'])
    if c.get('rotate_out_of_loop', False):
        coord = None
    return inp, cfg, coord, reg


def run_oracle_case(name, dtype=torch.float32):
    inp, cfg, coord, reg = case_cfg(name)
    c = inp['cfg']
    want = ['obj']
    if c.get('optimize_probe', False):
        want += ['probe_real', 'probe_imag']
    ppc = None
    if inp['probe_pos_correction'] is not None:
        ppc = torch.tensor(inp['probe_pos_correction']).to(dtype)
        want += ['probe_pos_correction']
    if reg is not None and reg.get('weight_l1') is not None:
        reg = dict(reg, weight_l1=torch.tensor(reg['weight_l1']).to(dtype))
    extra = {}
    if c.get('optimize_probe_pos_offset', False):
        extra['probe_pos_offset'] = torch.tensor(np.asarray(c['probe_pos_offset'], dtype='float32'))
        want += ['probe_pos_offset']
    if c.get('optimize_prj_pos_offset', False):
        extra['prj_pos_offset'] = torch.tensor(np.asarray(c['prj_pos_offset'], dtype='float32'))
        want += ['prj_pos_offset']
    if c.get('optimize_slice_pos', False):
        want += ['slice_pos_cm_ls']
    if c.get('optimize_tilt', False):
        extra['tilt'] = torch.tensor(np.asarray(c['tilt'], dtype='float32'))       # tilt_ls[:, i_theta], fp32 like the driver
        want += ['tilt']
    loss, data_loss, pred, grads = orc.ptycho_loss_and_grads(
        torch.tensor(inp['obj']).to(dtype), torch.tensor(inp['probe_real']).to(dtype),
        torch.tensor(inp['probe_imag']).to(dtype), inp['pos'], torch.tensor(inp['prj']).to(dtype), coord, cfg, reg=reg,
        probe_pos_correction=ppc, this_i_theta=c['i_theta'], this_ind_batch=inp['this_ind_batch'], want=tuple(want), **extra)
    return inp, float(loss), pred.numpy(), {k: v.numpy() for k, v in grads.items()}


def md_case_cfg(name):
    from cases import make_md_inputs
    inp = make_md_inputs(name)
    c = inp['cfg']
    N = c['N']
    cfg = dict(this_obj_size=[N, N, 1], probe_size=[N, N], energy_ev=c['energy_ev'], psize_cm=c['psize_cm'], h=None,
               unknown_type=c['unknown_type'], binning=1, sign_convention=1, scale_ri_by_k=True, fresnel_approx=True,
               two_d_mode=True, loss_function_type=c['loss'], raw_data_type=c['raw_data_type'])
    return inp, cfg


def run_md_oracle_case(name, dtype=torch.float32):
    """MultiDistModel loss + gradients from the oracle (oracle.multidist_loss_and_grads)."""
    inp, cfg = md_case_cfg(name)
    c = inp['cfg']
    want = ['obj']
    if c['optimize_probe']:
        want += ['probe_real', 'probe_imag']
    ppc = None
    if inp['probe_pos_correction'] is not None:
        ppc = torch.tensor(inp['probe_pos_correction']).to(dtype)
        want += ['probe_pos_correction']
    if c.get('free_prop', False):
        want += ['free_prop_cm']
    tiles = None
    if c.get('tiles'):
        from cases import md_tile_positions
        t = c['tiles']
        tiles = dict(pos_batch=md_tile_positions(c['N'], t['sub'])[np.array(t['ind'])], subprobe_size=[t['sub'], t['sub']],
                     safe_zone_width=t['szw'], n_dp_batch=t['n_dp_batch'])
        nb = (c['N'] // t['sub']) ** 2
        sel = np.concatenate([np.array(t['ind']) + d * nb for d in range(len(c['dists']))])
        inp = dict(inp, prj=inp['prj'][sel])
    aff = None
    if c.get('affine', False):
        aff = torch.tensor(inp['prj_affine_ls']).to(dtype)
        want += ['prj_affine_ls']
    loss, pred, grads = orc.multidist_loss_and_grads(
        torch.tensor(inp['obj']).to(dtype), torch.tensor(inp['probe_real']).to(dtype),
        torch.tensor(inp['probe_imag']).to(dtype), c['dists'], torch.tensor(inp['prj']).to(dtype), cfg,
        probe_pos_correction=ppc, want=tuple(want), prj_affine_ls=aff, tiles=tiles)
    return inp, float(loss), pred.numpy(), {k: v.numpy() for k, v in grads.items()}


def check_md_against_golden(gold, loss, pred, grads, tol=1e-4, gtol=1e-4):
    """Compare a MultiDist result (full arrays) with the compact golden of make_golden_multidist.py."""
    assert abs(loss - float(gold['loss'])) <= tol * abs(float(gold['loss']))
    assert rel_err(pred[:, ::4, ::4], gold['pred_sub']) < tol
    cg = compact_grad(grads['obj'])
    assert abs(cg['norm'] - float(gold['gobj_norm'])) <= gtol * float(gold['gobj_norm'])
    for k in ('proj0', 'proj1', 'proj2', 'sample'):
        assert rel_err(cg[k], gold['gobj_' + k]) < gtol, k
    for k in ('probe_real', 'probe_imag'):
        if 'g_' + k + '_sub' in gold.files:
            assert rel_err(grads[k][:, ::4, ::4], gold['g_' + k + '_sub']) < gtol, k
    if 'g_probe_pos_correction' in gold.files:
        assert rel_err(grads['probe_pos_correction'], gold['g_probe_pos_correction']) < gtol
    if 'g_free_prop_cm' in gold.files:
        assert rel_err(grads['free_prop_cm'], gold['g_free_prop_cm']) < gtol
    if 'g_prj_affine_ls' in gold.files:
        assert rel_err(grads['prj_affine_ls'], gold['g_prj_affine_ls']) < 5 * gtol


def check_param_grads_against_golden(gold, grads, i_theta, tol=1e-4):
    """Refinable-parameter gradients the goldens hold as full per-angle tables: tilt_ls [3, n_theta] and
    probe_pos_offset [n_theta, 2]; `grads` holds this angle's entries ('tilt' [3], 'probe_pos_offset' [2])."""
    if 'g_tilt_ls' in gold.files:
        assert rel_err(grads['tilt'], gold['g_tilt_ls'][:, i_theta]) < tol, 'tilt'
    if 'g_probe_pos_offset' in gold.files:
        assert rel_err(grads['probe_pos_offset'], gold['g_probe_pos_offset'][i_theta]) < tol, 'probe_pos_offset'
    if 'g_slice_pos_cm_ls' in gold.files:
        assert rel_err(grads['slice_pos_cm_ls'], gold['g_slice_pos_cm_ls']) < tol, 'slice_pos_cm_ls'
    if 'g_prj_pos_offset' in gold.files:
        assert rel_err(grads['prj_pos_offset'], gold['g_prj_pos_offset'][i_theta]) < tol, 'prj_pos_offset'
