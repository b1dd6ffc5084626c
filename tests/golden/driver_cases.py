"""Seeded inputs of the driver goldens (shared by make_golden_driver.py, the oracle driver test and the GPU test):
a small 3-D ptychotomography reconstruction that switches on everything the batch loop of
adorym/ptychography.py:859-1271 does around the hot path -- probe refinement, probe-position refinement with drift
removal, finite-support mask, reweighted L1 + TV, non-negativity, per-angle Adam step index."""
import numpy as np
from cases import _smooth_noise, _gaussian_probe

DRIVER_CASES = {
    'driver_immediate': dict(update_scheme='immediate', n_epochs=2),
    'driver_per_angle': dict(update_scheme='per angle', n_epochs=2),
    # rotate_out_of_loop (ptychography.py:918-943,1011,1067-1078): the object is rotated once per angle, the minibatches work
    # on obj.arr_rot, the gradient is rotated back with the inverse-angle table after every minibatch
    'driver_rot_out_of_loop': dict(update_scheme='immediate', n_epochs=1, rotate_out_of_loop=True),
    # tilt refinement (three differentiable rotations per angle, ptychography.py:725-726,919-930, forward_model.py:274-276)
    # + per-angle probe position offset (optimizers.py:845-861); checked against the reference golden on the GPU only
    # (the oracle driver does not restate the parameter plumbing of these two)
    'driver_tilt': dict(update_scheme='immediate', n_epochs=1, gpu_only=True,
                        extra=dict(optimize_tilt=True, tilt_learning_rate=2e-3, optimize_probe_pos_offset=True,
                                   probe_pos_offset_learning_rate=1e-2, optimize_all_probe_pos=False)),
    # plain gradient descent with the step-halving schedule (optimizers.py:432-464), phase-only object (beta pinned to 0,
    # ptychography.py:1138-1142), finite-support mask SHRUNK every second batch (shrink_cycle, :1216-1220), TV only
    'driver_gd_phase_only': dict(update_scheme='immediate', n_epochs=2, gpu_only=True,
                                 extra=dict(optimizer='gd', learning_rate=3e-6, object_type='phase_only', shrink_cycle=2,
                                            shrink_threshold=1e-4, optimize_probe=False, optimize_all_probe_pos=False,
                                            reweighted_l1=False, alpha_d=0, alpha_b=0, non_negativity=False)),
    # momentum optimizer (optimizers.py:366-411), Poisson loss on the magnitude data (forward_model.py:96-102), absorption-only
    # object (delta pinned to 0), update per angle
    'driver_momentum_poisson': dict(update_scheme='per angle', n_epochs=2, gpu_only=True,
                                    extra=dict(optimizer='momentum', learning_rate=2e-7, object_type='absorption_only',
                                               loss_function_type='poisson', optimize_probe=False,
                                               optimize_all_probe_pos=False, reweighted_l1=False, non_negativity=True)),
    # 2-D ptychography, the shape of BASELINE configs B / C scaled down: real_imag object, intensity data, two supplied probe
    # modes, probe + position refinement, minibatches of 3 of 9 positions (randomised order), 3 epochs
    'driver_2d_realimag': dict(kind='2d', update_scheme='immediate', n_epochs=3, gpu_only=True),
}
OBJ_SIZE = (32, 40, 40)
P = 72
N_THETA = 3
POS = [(-20.3, -18.6), (-24.8, -6.1), (-16.2, 1.4), (-9.7, -21.2), (-30.4, -30.9), (-12.5, -11.3)]


def driver_params(name):
    c = DRIVER_CASES[name]
    rs = np.random.RandomState(77)
    init = [_smooth_noise(rs, OBJ_SIZE, 3e-4), _smooth_noise(rs, OBJ_SIZE, 3e-5)]
    p = dict(obj_size=OBJ_SIZE, probe_pos=POS, theta_st=0, theta_end=2 * np.pi, n_theta=N_THETA, energy_ev=5000.,
                psize_cm=1e-7, free_prop_cm='inf', minibatch_size=2, n_epochs=c['n_epochs'], alpha_d=2e-2, alpha_b=2e-3,
                gamma=5e-2, reweighted_l1=True, learning_rate=2e-5, optimizer='adam', update_scheme=c['update_scheme'],
                probe_type='gaussian', probe_mag_sigma=6, probe_phase_sigma=6, probe_phase_max=0.5,
                optimize_probe=True, probe_learning_rate=1e-3, optimize_all_probe_pos=True, all_probe_pos_learning_rate=1e-2,
                non_negativity=True, initial_guess=init, unknown_type='delta_beta', loss_function_type='lsq',
                raw_data_type='magnitude', randomize_probe_pos=False, shared_probe_among_angles=True,
                rotate_out_of_loop=c.get('rotate_out_of_loop', False))
    p.update(c.get('extra', {}))
    return p


def driver_truth():
    rs = np.random.RandomState(78)
    return np.stack([_smooth_noise(rs, OBJ_SIZE, 6e-4), _smooth_noise(rs, OBJ_SIZE, 6e-5)], -1)


def driver_mask():
    Y, X, Z = OBJ_SIZE
    yy, xx, zz = np.meshgrid(np.arange(Y) - Y / 2 + 0.5, np.arange(X) - X / 2 + 0.5, np.arange(Z) - Z / 2 + 0.5, indexing='ij')
    return ((xx ** 2 + zz ** 2) < (0.45 * X) ** 2).astype('float32')


def driver_data():
    """Measured magnitudes [n_theta, n_pos, P, P] (complex64, as exchange/data stores them): the oracle's forward model on
    the true object at the ROUNDED positions with the initial Gaussian probe (adorym/simulation.py:659-671 recipe)."""
    import torch
    from oracle import multislice_oracle as orc
    Y, X, Z = OBJ_SIZE
    theta_ls = np.linspace(0, 2 * np.pi, N_THETA, dtype='float32')
    voxel_nm = np.array([1e-7] * 3) * 1e7
    h = orc.get_kernel(voxel_nm[-1], 1240. / 5000., voxel_nm, [P, P])
    cfg = dict(this_obj_size=[Y, X, Z], probe_size=[P, P], energy_ev=5000., psize_cm=1e-7, h=h, free_prop_cm='inf',
               unknown_type='delta_beta')
    pr, pi = orc.gaussian_probe([P, P], 1, 6, 0.5, 6)
    pr = torch.tensor(pr[None], dtype=torch.float32) * 1.1      # the "true" probe differs from the initial guess
    pi = torch.tensor(pi[None], dtype=torch.float32) * 0.9
    pos = np.round(np.array(POS)).astype(int)
    tr = torch.tensor(driver_truth(), dtype=torch.float32)
    data = np.zeros([N_THETA, len(pos), P, P], dtype='complex64')
    with torch.no_grad():
        for i in range(N_THETA):
            data[i] = orc.predict_ptycho(tr, pr, pi, pos, orc.rotation_coords([Y, X, Z], theta_ls[i]), cfg).numpy()
    return data


# ---------------------------------------------------------------------------------------------------------------
# 2-D ptychography driver case
# ---------------------------------------------------------------------------------------------------------------
OBJ2D = (100, 92, 1)
P2D = 64
POS2D = [(-8.2, -7.6), (9.7, 12.1), (30.4, 20.2), (34.6, -4.9), (2.8, 27.3), (18.1, 3.3), (25.5, 28.4), (-3.1, 11.8), (12.4, -6.2)]


def probe_modes_2d():
    """Two probe modes as [mag, phase] stacks [M, P, P] (probe_type='supplied', ptychography.py:105,234-235)."""
    rs = np.random.RandomState(81)
    r0, i0 = _gaussian_probe((P2D, P2D), 7.0, 0.4, 7.0)
    pr = np.stack([r0 * (1 - 0.5 * m) + 0.03 * rs.normal(size=(P2D, P2D)) * np.abs(r0) for m in range(2)])
    pi = np.stack([i0 * (1 + 0.3 * m) + 0.03 * rs.normal(size=(P2D, P2D)) * np.abs(r0) for m in range(2)])
    c = pr + 1j * pi
    return np.abs(c).astype('float32'), np.angle(c).astype('float32')


def driver_params_2d(name):
    c = DRIVER_CASES[name]
    rs = np.random.RandomState(82)
    init = [(1.0 - _smooth_noise(rs, OBJ2D, 0.05)).astype('float32'), _smooth_noise(rs, OBJ2D, 0.05)]        # [mag, phase]
    mag, ph = probe_modes_2d()
    return dict(obj_size=OBJ2D, probe_pos=POS2D, theta_st=0, theta_end=0, n_theta=1, energy_ev=8801.12, psize_cm=1.3279e-6,
                free_prop_cm='inf', minibatch_size=3, n_epochs=c['n_epochs'], alpha_d=0, alpha_b=0, gamma=0,
                learning_rate=1e-3, optimizer='adam', update_scheme=c['update_scheme'], two_d_mode=True,
                probe_type='supplied', probe_initial=[mag, ph], n_probe_modes=2, optimize_probe=True, probe_learning_rate=1e-3,
                optimize_all_probe_pos=True, all_probe_pos_learning_rate=1e-2, initial_guess=init, unknown_type='real_imag',
                loss_function_type='lsq', raw_data_type='intensity', randomize_probe_pos=True, shared_probe_among_angles=True)


def driver_data_2d():
    """Intensities [1, n_pos, P, P] from the oracle's forward model of a 'true' real_imag object with perturbed probe modes."""
    import torch
    from oracle import multislice_oracle as orc
    rs = np.random.RandomState(83)
    mag = 1.0 - _smooth_noise(rs, OBJ2D, 0.35)
    ph = _smooth_noise(rs, OBJ2D, 0.8)
    tr = torch.tensor(np.stack([mag * np.cos(ph), mag * np.sin(ph)], -1).astype('float32'))
    pm, pp = probe_modes_2d()
    pr = torch.tensor(pm * np.cos(pp) * 1.08)
    pi = torch.tensor(pm * np.sin(pp) * 0.93)
    cfg = dict(this_obj_size=list(OBJ2D), probe_size=[P2D, P2D], energy_ev=8801.12, psize_cm=1.3279e-6, h=None,
               free_prop_cm='inf', unknown_type='real_imag', two_d_mode=True)
    pos = np.round(np.array(POS2D)).astype(int)
    with torch.no_grad():
        a = orc.predict_ptycho(tr, pr, pi, pos, None, cfg).numpy()
    return (a ** 2)[None].astype('float32')
