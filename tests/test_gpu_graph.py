"""GPU: the CUDA-graph replay of a minibatch (adorym_b200.graph.GraphedDifferentiator) against the eager evaluation of the
same minibatches -- changing angle, patch positions, data indices, with in-place Adam updates of the object / probe /
position corrections in between -- and through the driver (use_cuda_graph=True) against the reference driver golden."""
import os
import numpy as np
import pytest
import torch

from cases import CASES
from oracle_util import case_cfg, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _setup(name):
    import adorym_b200 as ab
    from test_gpu_parity import _common_vars
    dev = torch.device('cuda', 0)
    inp, cfg, coord, reg = case_cfg(name)
    c = inp['cfg']
    n_theta, P = c['n_theta'], c['probe_size']
    theta_ls = np.linspace(0, 2 * np.pi, n_theta, dtype='float32')
    rs = np.random.RandomState(3)
    prj_full = torch.tensor(rs.uniform(0.2, 1.0, size=(n_theta, 16, P[0], P[1])).astype('float32') * float(inp['prj'].mean()),
                            device=dev)
    cv = _common_vars(inp, cfg, dev, prj_full, theta_ls, n_theta)
    fm = ab.PtychographyModel(loss_function_type=c['loss'], device=dev, common_vars_dict=cv, raw_data_type=c['raw_data_type'])
    if c.get('alpha_d'):
        fm.add_regularizers([ab.L1Regularizer(c['alpha_d'], c['alpha_b']), ab.TVRegularizer(c['gamma'])])
    return ab, dev, inp, c, fm, prj_full, n_theta


@pytest.mark.parametrize('name', ['ptycho3d_p72', 'ptycho3d_p72_edge', 'ptycho2d_modes'])
def test_graph_replay_equals_eager(name):
    ab, dev, inp, c, fm, prj_full, n_theta = _setup(name)
    t = lambda a, rg=True: torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=rg)
    refine = c.get('optimize_probe', False)
    has_ppc = inp['probe_pos_correction'] is not None
    B = len(inp['pos'])
    state = dict(obj=t(inp['obj']), pr=t(inp['probe_real'], refine), pi=t(inp['probe_imag'], refine),
                 ppc=t(inp['probe_pos_correction']) if has_ppc else t(np.zeros([n_theta, 16, 2]), False))
    opt_args = [0] + ([fm.get_argument_index('probe_real'), fm.get_argument_index('probe_imag')] if refine else []) \
        + ([fm.get_argument_index('probe_pos_correction')] if has_ppc else [])
    eager, graphed = ab.Differentiator(), ab.GraphedDifferentiator(fm)
    eager.create_loss_node(fm.get_loss_function(), opt_args)
    graphed.create_loss_node(fm.get_loss_function(), opt_args)
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': 1e-5})
    opt.create_container(list(state['obj'].shape), False, dev)
    rs = np.random.RandomState(5)
    for step in range(4):
        it = int(rs.randint(0, n_theta)) if c['obj_size'][2] > 1 else 0
        pos = inp['pos'] + rs.randint(-3, 4, size=inp['pos'].shape)
        ind = np.sort(rs.choice(16, size=B, replace=False))
        args = dict(obj=state['obj'], probe_real=state['pr'], probe_imag=state['pi'], probe_defocus_mm=None,
                    probe_pos_offset=None, this_i_theta=it, this_pos_batch=pos, prj=prj_full,
                    probe_pos_correction=state['ppc'], this_ind_batch=ind, tilt_ls=None, prj_pos_offset=None)
        ge = [g.clone() for g in eager.get_gradients(**args)]
        le = fm.current_loss
        gg = graphed.get_gradients(**args)
        lg = fm.current_loss
        assert graphed.fallback_reason is None, graphed.fallback_reason
        assert abs(lg - le) <= 1e-6 * abs(le)
        for a, b in zip(gg, ge):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
        with torch.no_grad():                                   # in-place update: the graph's input storage stays put
            state['obj'] = opt.apply_gradient(state['obj'], ge[0], step, step_size=1e-5).requires_grad_(True)
            if has_ppc:
                state['ppc'] = (state['ppc'].detach() - 0.05 * ge[-1]).requires_grad_(True)     # NEW storage: copied into the static
    assert graphed.n_replays == 4 and len(graphed._graphs) == 1


def test_driver_with_cuda_graph_matches_reference_driver_golden():
    """reconstruct_ptychography(use_cuda_graph=True) on the driver golden of the unmodified reference (probe + position
    refinement, mask, reweighted L1 + TV, non-negativity, 2 epochs): same tolerances as the eager driver test."""
    import adorym_b200 as ab
    from driver_cases import driver_params, driver_mask
    from test_driver_golden import _check, _data
    for name in ('driver_immediate', 'driver_per_angle'):
        out = ab.reconstruct_ptychography(_data(), **driver_params(name), finite_support_mask_path=driver_mask(), verbose=False,
                                          save_outputs=False, output_folder='/tmp/adm_test_driver_graph_' + name,
                                          use_cuda_graph=True)
        hist = [(e, b, l) for (e, b, l, t) in out['loss_history']]
        _check(name, hist, out['obj'].cpu().numpy(), out['probe_real'].cpu().numpy(), out['probe_imag'].cpu().numpy(),
               out['probe_pos_correction'].cpu().numpy())


def test_graph_replay_equals_eager_multidist():
    """MultiDistModel (4 distances, refinable data shift) replayed from a CUDA graph vs eager, with updates in between."""
    import adorym_b200 as ab
    from cases import make_md_inputs
    from test_gpu_multidist import md_common_vars
    dev = torch.device('cuda', 0)
    inp = make_md_inputs('multidist256_shift')
    c = inp['cfg']
    prj_full = torch.tensor(inp['prj'][None], device=dev)
    fm = ab.MultiDistModel(loss_function_type=c['loss'], distribution_mode=None, device=dev,
                           common_vars_dict=md_common_vars(c, dev, prj_full, True), raw_data_type=c['raw_data_type'])
    t = lambda a, rg=True: torch.tensor(a, dtype=torch.float32, device=dev, requires_grad=rg)
    state = dict(obj=t(inp['obj']), ppc=t(inp['probe_pos_correction']))
    pr, pi = t(inp['probe_real'], False), t(inp['probe_imag'], False)
    opt_args = [0, fm.get_argument_index('probe_pos_correction')]
    eager, graphed = ab.Differentiator(), ab.GraphedDifferentiator(fm)
    eager.create_loss_node(fm.get_loss_function(), opt_args)
    graphed.create_loss_node(fm.get_loss_function(), opt_args)
    for step in range(3):
        args = dict(obj=state['obj'], probe_real=pr, probe_imag=pi, probe_defocus_mm=None, probe_pos_offset=None, this_i_theta=0,
                    this_pos_batch=np.array([[0, 0]]), prj=prj_full, probe_pos_correction=state['ppc'],
                    this_ind_batch=np.array([0]), free_prop_cm=c['dists'], safe_zone_width=0, prj_affine_ls=None,
                    ctf_lg_kappa=None, prj_pos_offset=None)
        ge = [g.clone() for g in eager.get_gradients(**args)]
        le = fm.current_loss
        gg = graphed.get_gradients(**args)
        assert graphed.fallback_reason is None, graphed.fallback_reason
        assert abs(fm.current_loss - le) <= 1e-6 * abs(le)
        for a, b in zip(gg, ge):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
        with torch.no_grad():
            state['obj'] = (state['obj'].detach() - 0.01 * ge[0]).requires_grad_(True)
            state['ppc'] = (state['ppc'].detach() - 0.05 * ge[1]).requires_grad_(True)
    assert graphed.n_replays == 3 and len(graphed._graphs) == 1
