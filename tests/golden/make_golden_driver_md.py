"""Golden of the DRIVER on multi-distance holography: the UNMODIFIED reference's reconstruct_ptychography (PyTorch backend,
CPU, I/O shims) with MultiDistModel, refinable propagation distances AND refinable affine alignment of the holograms -- the
option set of demos/2d_multidist_holography_w_affine.py on 128^2 frames, 3 distances, 5 epochs (one update each).  Stored:
the measured intensities (fp16), the loss history, the final object, distances and affine matrices.  Build container only:

    PYTHONPATH=oracle/ref_shims:/root/reference:. python tests/golden/make_golden_driver_md.py
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import _smooth_noise  # noqa: E402

N = 128
TRUE_DISTS = [35., 62., 90.]
START_DISTS = [35., 60., 93.]
ENERGY_EV, PSIZE_CM = 17500., 1e-4


def md_driver_params():
    rs = np.random.RandomState(91)
    init = [(1.0 - _smooth_noise(rs, (N, N, 1), 0.05)).astype('float32'), _smooth_noise(rs, (N, N, 1), 0.05)]
    return dict(obj_size=(N, N, 1), probe_pos=[(0, 0)], theta_st=0, theta_end=0, n_theta=1, energy_ev=ENERGY_EV,
                psize_cm=PSIZE_CM, free_prop_cm=list(START_DISTS), minibatch_size=1, n_epochs=5, alpha_d=0, alpha_b=0, gamma=0,
                learning_rate=1e-2, optimizer='adam', update_scheme='immediate', probe_type='plane', two_d_mode=True,
                optimize_probe=False, optimize_all_probe_pos=False, optimize_free_prop=True, free_prop_learning_rate=1e-1,
                optimize_prj_affine=True, prj_affine_learning_rate=1e-3, initial_guess=init, unknown_type='real_imag',
                loss_function_type='lsq', raw_data_type='intensity', randomize_probe_pos=False, safe_zone_width=0)


def md_driver_data():
    """Intensities [1, n_dists, N, N]: the oracle's MultiDist forward model of a "true" object at the true distances, the
    holograms of distances 1 and 2 then resampled by small affine distortions (what the refinement has to undo)."""
    import torch
    from oracle import multislice_oracle as orc
    rs = np.random.RandomState(92)
    mag = 1.0 - _smooth_noise(rs, (N, N, 1), 0.4)
    ph = _smooth_noise(rs, (N, N, 1), 0.9)
    obj = torch.tensor(np.stack([mag * np.cos(ph), mag * np.sin(ph)], -1).astype('float32'))
    cfg = dict(this_obj_size=[N, N, 1], probe_size=[N, N], energy_ev=ENERGY_EV, psize_cm=PSIZE_CM, h=None,
               unknown_type='real_imag', two_d_mode=True)
    with torch.no_grad():
        pred = orc.predict_multidist(obj, torch.ones(1, N, N), torch.zeros(1, N, N), TRUE_DISTS, cfg)
        inten = pred ** 2
        mats = torch.tensor([[[1., 0, 0], [0, 1., 0]], [[1.01, 0.004, 0.01], [-0.003, 0.99, -0.008]],
                             [[0.985, -0.005, -0.012], [0.006, 1.012, 0.006]]])
        frames = torch.cat([orc.affine_transform(inten[i:i + 1], mats[i]) for i in range(3)])
    return frames.numpy()[None].astype('float32')


def run(data):
    import adorym  # noqa: F401
    import adorym.ptychography as ptych
    import adorym.optimizers as ropt
    last = {}
    orig = ropt.update_parameters

    def spy(opt_ls, optimizable_params, kwargs):      # records what the reference's own function returns; no arithmetic
        out = orig(opt_ls, optimizable_params, kwargs)
        last.update({k: v for k, v in out.items()})
        return out
    ptych.update_parameters = spy
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        np.savez(os.path.join(tmp, 'data.h5'), **{'exchange|data': data})
        os.rename(os.path.join(tmp, 'data.h5.npz'), os.path.join(tmp, 'data.h5'))
        params = md_driver_params()
        # the reference's driver hands run_bfloat16 / run_float64 to every forward model, and its MultiDistModel forwards
        # them to ForwardModel.__init__, which rejects them (ptychography.py:526-535, forward_model.py:811-814): with
        # forward_model='auto' the unmodified reference cannot construct the class.  The documented `forward_model`
        # argument takes a user-supplied factory, which drops the two flags; nothing else differs from 'auto'.
        from adorym.forward_model import MultiDistModel

        def factory(**kw):
            kw.pop('run_bfloat16', None)
            kw.pop('run_float64', None)
            return MultiDistModel(**kw)
        params['forward_model'] = factory
        params.update(fname='data.h5', save_path=tmp, output_folder=os.path.join(tmp, 'out'), backend='pytorch', cpu_only=True,
                      store_checkpoint=False, use_checkpoint=False, save_intermediate=False, n_batch_per_checkpoint=1000)
        ptych.reconstruct_ptychography(**params)
        hist = np.loadtxt(os.path.join(tmp, 'out', 'convergence', 'loss_rank_0.txt'), delimiter=',', ndmin=2, skiprows=1)
        names = sorted(os.listdir(os.path.join(tmp, 'out')))
        re_f = [n for n in names if n.endswith('.npy')]
        arrs = {n[:-4]: np.load(os.path.join(tmp, 'out', n)) for n in re_f}
        out = {'loss_history': hist[:, :3].astype('float64'),
               'free_prop_cm': last['free_prop_cm'].detach().numpy().astype('float32'),
               'prj_affine_ls': last['prj_affine_ls'].detach().numpy().astype('float32')}
        for k, v in arrs.items():
            out['out_' + k] = v.astype('float32')
        return out
    finally:
        ptych.update_parameters = orig
        os.chdir(cwd)


if __name__ == '__main__':
    data = md_driver_data().astype('float16').astype('float32')       # stored compactly; every consumer uses these values
    out = run(data)
    out['data'] = data.astype('float16')
    path = os.path.join(HERE, 'driver_md_affine.npz')
    np.savez_compressed(path, **out)
    print({k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})
    print('losses', out['loss_history'][:, 2], out['free_prop_cm'], out['prj_affine_ls'], os.path.getsize(path) // 1024, 'KB')
