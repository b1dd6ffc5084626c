"""Golden of the reference's OWN CI test (tests/test_multislice_tomography_64.py: 64^3 full-field multislice tomography of
the shipped adhesin data set, SingleBatchFullfieldModel, L1 regulariser, finite-support mask, Adam 1e-7, one epoch) run
with the UNMODIFIED reference (PyTorch backend on CPU, I/O shims).  Changes to the test's parameters: a seeded
`initial_guess` instead of the time-seeded random one, file outputs off, and theta_downsample = 50 instead
of 10 (10 of the 500 angles = 10 minibatches, so that the fp32 frames the run reads fit in a small fixture; the magnitudes
are 1 - O(1e-5), so they cannot be stored in fp16).  Stored: those 10 measured frames (fp32 magnitudes), the mask
(bit-packed), the loss history and the final object, plus the same run repeated with the initial guess perturbed by 1e-7
relative (the reference's own sensitivity, which sets the tolerance of the test).  Run in the build container only:

    PYTHONPATH=oracle/ref_shims:/root/reference:. python tests/golden/make_golden_ci.py
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DEMO = '/root/reference/demos/adhesin'


DS = 50


def ci_initial_guess():
    """Seeded version of util.py:71-124 for delta_beta objects: N(8.7e-7, 1e-7) / N(5.1e-8, 1e-8), clipped at 0."""
    rs = np.random.RandomState(1234)
    d = rs.normal(8.7e-7, 1e-7, size=(64, 64, 64)).astype('float32')
    b = rs.normal(5.1e-8, 1e-8, size=(64, 64, 64)).astype('float32')
    return [np.clip(d, 0, None), np.clip(b, 0, None)]


def ci_params():
    return dict(theta_st=0, theta_end=2 * np.pi, theta_downsample=DS, n_epochs=1, obj_size=(64, 64, 64),
                learning_rate=1e-7, energy_ev=800, psize_cm=0.67e-7, binning=1, minibatch_size=1, n_batch_per_update=1,
                multiscale_level=1, n_epoch_final_pass=None, free_prop_cm=0, initial_guess=ci_initial_guess(),
                probe_initial=None, n_dp_batch=1, fresnel_approx=True, probe_type='plane', forward_algorithm='fresnel',
                object_type='normal', probe_pos=[(0, 0)], optimize_probe_defocusing=False, distribution_mode=None,
                optimizer='adam', use_checkpoint=False, reweighted_l1=True)


def _run(data, mask, perturb):
    import adorym
    from adorym.ptychography import reconstruct_ptychography
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        np.savez(os.path.join(tmp, 'data.h5'), **{'exchange|data': data})
        os.rename(os.path.join(tmp, 'data.h5.npz'), os.path.join(tmp, 'data.h5'))
        np.save(os.path.join(tmp, 'mask.npy'), mask)
        params = ci_params()
        if perturb:
            rs = np.random.RandomState(5)
            params['initial_guess'] = [(a * (1 + perturb * rs.normal(size=a.shape))).astype('float32') for a in params['initial_guess']]
        params.update(fname='data.h5', save_path=tmp, output_folder=os.path.join(tmp, 'out'), cpu_only=True, backend='pytorch',
                      regularizers=[adorym.L1Regularizer(alpha_d=1.e-9 * 64 ** 3, alpha_b=1.e-10 * 64 ** 3)],
                      finite_support_mask_path=os.path.join(tmp, 'mask.npy'), save_intermediate=False, store_checkpoint=False)
        reconstruct_ptychography(**params)
        hist = np.loadtxt(os.path.join(tmp, 'out', 'convergence', 'loss_rank_0.txt'), delimiter=',', ndmin=2, skiprows=1)
        obj = np.stack([np.load(os.path.join(tmp, 'out', 'delta_ds_1.npy')), np.load(os.path.join(tmp, 'out', 'beta_ds_1.npy'))], -1)
        return hist[:, :3].astype('float64'), obj.astype('float32')
    finally:
        os.chdir(cwd)


def main():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle', 'ref_shims'))
    import h5py     # the shim: np.memmap over the shipped file
    import dxchange
    f = h5py.File(os.path.join(REF_DEMO, 'data_adhesin_360_soft_4d.h5'), 'r')
    full = f['exchange/data']
    used = np.abs(np.asarray(full[0:500:DS])).astype('float32')             # the frames the run reads
    mask = np.asarray(dxchange.read_tiff(os.path.join(REF_DEMO, 'fin_sup_mask', 'mask.tiff'))).astype('float32')
    data = np.zeros((500, 1, 64, 64), dtype='complex64')
    data[0:500:DS] = used
    hist, obj = _run(data, mask, 0.0)
    hist_p, obj_p = _run(data, mask, 1e-7)
    path = os.path.join(HERE, 'driver_ci64.npz')
    sens_l = np.max(np.abs(hist_p[:, 2] - hist[:, 2]) / hist[:, 2])
    sens_o = np.linalg.norm(obj_p - obj) / np.linalg.norm(obj)
    np.savez_compressed(path, frames=used, mask_bits=np.packbits(mask > 0.5), loss_history=hist, obj=obj,
                        sensitivity_loss_history=np.float64(sens_l), sensitivity_obj=np.float64(sens_o))
    print('driver_ci64 losses', hist[:3, 2], '...', hist[-1, 2], '->', os.path.getsize(path) // 1024, 'KB; mask binary:',
          bool(np.all((mask == 0) | (mask == 1))), '; sensitivity to a 1e-7 perturbation: loss history %.2e, object %.2e' % (sens_l, sens_o))


if __name__ == '__main__':
    main()
