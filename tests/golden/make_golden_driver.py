"""Golden of the DRIVER: the UNMODIFIED reference's reconstruct_ptychography (adorym/ptychography.py:54-1296, PyTorch
backend on CPU, through the I/O shims) on the small synthetic reconstruction of driver_cases.py -- 2 epochs x 3 angles x
3 minibatches with probe + probe-position refinement, finite-support mask, reweighted L1 + TV, non-negativity,
update_scheme 'immediate' and 'per angle'.  Stored: the measured data, the per-batch loss history
(convergence/loss_rank_0.txt), the final object, probe and probe-position corrections.  Run in the build container only:

    PYTHONPATH=oracle/ref_shims:/root/reference:. python tests/golden/make_golden_driver.py
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from driver_cases import (DRIVER_CASES, driver_params, driver_data, driver_mask, driver_params_2d,  # noqa: E402
                          driver_data_2d)


def run(name, data):
    import adorym
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
        two_d = DRIVER_CASES[name].get('kind') == '2d'
        if two_d:
            params = driver_params_2d(name)
            params.update(fname='data.h5', save_path=tmp, output_folder=os.path.join(tmp, 'out'), backend='pytorch', cpu_only=True,
                          store_checkpoint=False, use_checkpoint=False, save_intermediate=False, n_batch_per_checkpoint=1000)
        else:
            np.save(os.path.join(tmp, 'mask.npy'), driver_mask())
            params = driver_params(name)
            params.update(fname='data.h5', save_path=tmp, output_folder=os.path.join(tmp, 'out'), backend='pytorch', cpu_only=True,
                          finite_support_mask_path=os.path.join(tmp, 'mask.npy'), store_checkpoint=False, use_checkpoint=False,
                          save_intermediate=False, n_batch_per_checkpoint=1000)
        ptych.reconstruct_ptychography(**params)
        hist = np.loadtxt(os.path.join(tmp, 'out', 'convergence', 'loss_rank_0.txt'), delimiter=',', ndmin=2, skiprows=1)
        if two_d:
            # real_imag objects are written as magnitude / phase (ptychography.py:1275-1283)
            mag = np.load(os.path.join(tmp, 'out', 'obj_mag_ds_1.npy'))
            phs = np.load(os.path.join(tmp, 'out', 'obj_phase_ds_1.npy'))
            delta, beta = mag * np.cos(phs), mag * np.sin(phs)
        else:
            delta = np.load(os.path.join(tmp, 'out', 'delta_ds_1.npy'))
            beta = np.load(os.path.join(tmp, 'out', 'beta_ds_1.npy'))
        out = {'loss_history': hist[:, :3].astype('float64'), 'obj': np.stack([delta, beta], -1).astype('float32'),
               'probe_real': last['probe_real'].detach().numpy().astype('float32'),
               'probe_imag': last['probe_imag'].detach().numpy().astype('float32'),
               'probe_pos_correction': last['probe_pos_correction'].detach().numpy().astype('float32')}
        for k in ('tilt_ls', 'probe_pos_offset'):
            if k in last and last[k].requires_grad:
                out[k] = last[k].detach().numpy().astype('float32')
        return out
    finally:
        ptych.update_parameters = orig
        os.chdir(cwd)


def main():
    data = driver_data()
    if not sys.argv[1:]:
        np.savez_compressed(os.path.join(HERE, 'driver_data.npz'), data=np.abs(data).astype('float16'))
    # the data are stored as fp16 magnitudes (compact); every consumer (reference run included) uses the fp16-rounded values
    data = np.abs(data).astype('float16').astype('float32').astype('complex64')
    for name in (sys.argv[1:] or DRIVER_CASES):
        if DRIVER_CASES[name].get('kind') == '2d':
            d2 = driver_data_2d()                 # fp32 (far-field intensities exceed the fp16 range); stored in the golden itself
            out = run(name, d2)
            out['data'] = d2
            path = os.path.join(HERE, name + '.npz')
            np.savez_compressed(path, **out)
            print(name, 'losses', out['loss_history'][:3, 2], '...', out['loss_history'][-1, 2], '->', os.path.getsize(path) // 1024, 'KB')
            continue
        out = run(name, data)
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(name, 'losses', out['loss_history'][:3, 2], '...', out['loss_history'][-1, 2], '->', os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
