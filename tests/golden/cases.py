"""Seeded input generators for the golden/parity cases (shared by make_golden.py, the oracle tests
and the GPU parity tests).  numpy RandomState streams are stable across numpy versions."""
import numpy as np


def _gaussian_probe(size, mag_sigma, phase_max, phase_sigma):
    py = np.arange(size[0]) - (size[0] - 1.) / 2
    px = np.arange(size[1]) - (size[1] - 1.) / 2
    pxx, pyy = np.meshgrid(px, py)
    mag = np.exp(-(pxx ** 2 + pyy ** 2) / (2 * mag_sigma ** 2))
    ph = phase_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * phase_sigma ** 2))
    return mag * np.cos(ph), mag * np.sin(ph)


def _smooth_noise(rs, shape, scale):
    """Band-limited positive random field (sum of a few random cosines) * scale."""
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')
    f = np.zeros(shape)
    for _ in range(6):
        k = rs.uniform(-0.35, 0.35, size=len(shape))
        ph = rs.uniform(0, 2 * np.pi)
        f += np.cos(sum(2 * np.pi * ki * g for ki, g in zip(k, grids)) + ph)
    f = (f - f.min()) / (f.max() - f.min() + 1e-12)
    f = f + 0.15 * rs.uniform(size=shape)
    return (f * scale).astype('float32')


CASES = {
    # 3-D ptychotomography, the shape of BASELINE config D scaled down: P=72, overhanging probes.
    'ptycho3d_p72': dict(kind='ptycho', obj_size=(80, 80, 80), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                         theta=0.7, n_theta=7, i_theta=3,
                         pos=[(-36, -36), (4, 5), (-10, 30), (8, -20)], n_modes=1, free_prop_cm='inf',
                         unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=11,
                         loss='lsq', raw_data_type='magnitude'),
    # same geometry, Poisson loss on intensity data, sign_convention -1, normalize_fft
    'ptycho3d_p72_poisson': dict(kind='ptycho', obj_size=(80, 80, 80), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                                 theta=-1.9, n_theta=7, i_theta=5,
                                 pos=[(0, 0), (8, 8), (-5, 3)], n_modes=1, free_prop_cm='inf',
                                 unknown_type='delta_beta', delta_scale=5e-4, beta_scale=5e-5, seed=12,
                                 loss='poisson', raw_data_type='intensity', sign_convention=-1, normalize_fft=True,
                                 poisson_multiplier=3.0),
    # edge geometry: probe larger than the (non-cubic) object, a patch lying entirely outside it, odd and even columns,
    # an odd number of patterns, two probe modes with probe + position refinement through a 3-D object, L1 + TV
    'ptycho3d_p72_edge': dict(kind='ptycho', obj_size=(40, 48, 48), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                              theta=2.1, n_theta=5, i_theta=1,
                              pos=[(-20, -14), (-90, -90), (-36, 11), (3, -31), (-5, -6)], n_modes=2, free_prop_cm='inf',
                              unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=16,
                              loss='lsq', raw_data_type='magnitude', optimize_probe=True, optimize_all_probe_pos=True,
                              alpha_d=3e-3, alpha_b=3e-4, gamma=2e-3),
    # no detector transform (free_prop_cm = 0: the exit wave is detected), real_imag unknown, 3-D, P = 72
    'ptycho3d_p72_exitwave': dict(kind='ptycho', obj_size=(48, 48, 48), probe_size=(72, 72), energy_ev=5000.,
                                  psize_cm=1e-7, theta=-0.4, n_theta=4, i_theta=2,
                                  pos=[(-12, -12), (-30, 5), (1, -19)], n_modes=1, free_prop_cm=0,
                                  unknown_type='real_imag', delta_scale=0.2, beta_scale=0.2, seed=17,
                                  loss='lsq', raw_data_type='intensity'),
    # regulariser-dominated gradient: L1 + TV weights chosen so that |g_reg| is a large fraction of |g| (a wrong sign, a
    # missing TV neighbour or a wrong 1/N in the fused regulariser kernel cannot hide below the tolerance); non-cubic
    'ptycho3d_p72_strongreg': dict(kind='ptycho', obj_size=(40, 56, 48), probe_size=(72, 72), energy_ev=5000.,
                                   psize_cm=1e-7, theta=1.3, n_theta=5, i_theta=2,
                                   pos=[(-20, -14), (-30, 3), (3, -17)], n_modes=1, free_prop_cm='inf',
                                   unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=18,
                                   loss='lsq', raw_data_type='magnitude', alpha_d=4e4, alpha_b=2e4, gamma=3e4),
    # reweighted L1 (weights as the driver forms them, ptychography.py:995-999) + TV, also regulariser-dominated
    'ptycho3d_p72_rwl1': dict(kind='ptycho', obj_size=(40, 48, 56), probe_size=(72, 72), energy_ev=5000.,
                              psize_cm=1e-7, theta=-0.9, n_theta=5, i_theta=4,
                              pos=[(-16, -20), (-32, 0), (5, -11)], n_modes=1, free_prop_cm='inf',
                              unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=19,
                              loss='lsq', raw_data_type='magnitude', alpha_d=2e3, alpha_b=1e3, gamma=2e4,
                              reweighted=True),
    # beamstop: the central detector pixels are excluded from the loss (forward_model.py:128-136)
    'ptycho3d_p72_beamstop': dict(kind='ptycho', obj_size=(48, 48, 48), probe_size=(72, 72), energy_ev=5000.,
                                  psize_cm=1e-7, theta=0.5, n_theta=4, i_theta=1,
                                  pos=[(-12, -12), (-30, 5), (1, -19)], n_modes=2, free_prop_cm='inf',
                                  unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=20,
                                  loss='lsq', raw_data_type='magnitude', optimize_probe=True, beamstop_radius=7.5),
    # full-field multislice tomography on frames the shared-memory FFT cannot hold (128 pixels): large-frame row / column
    # kernels, one fused step per slice; exit wave (free_prop_cm = 0) and near-field detection
    'fullfield128': dict(kind='fullfield', obj_size=(128, 128, 10), probe_size=(128, 128), energy_ev=800., psize_cm=0.67e-7,
                         theta=0.0, n_theta=3, i_theta=0, pos=[(0, 0)], n_modes=1, free_prop_cm=0,
                         unknown_type='delta_beta', delta_scale=2e-3, beta_scale=3e-4, seed=24,
                         loss='lsq', raw_data_type='magnitude', rotate_out_of_loop=True),
    'fullfield128_near': dict(kind='fullfield', obj_size=(128, 128, 6), probe_size=(128, 128), energy_ev=800.,
                              psize_cm=0.67e-7, theta=0.0, n_theta=3, i_theta=0, pos=[(0, 0)], n_modes=1,
                              free_prop_cm=2e-5, unknown_type='real_imag', delta_scale=0.2, beta_scale=0.2, seed=25,
                              loss='lsq', raw_data_type='intensity', rotate_out_of_loop=True),
    # refinable shift of the exit wave (optimize_prj_pos_offset / shift_exit_wave, propagate.py:260-261: plane-wave
    # tomography alignment) followed by near-field detection
    'fullfield64_prjoffset': dict(kind='fullfield', obj_size=(64, 64, 12), probe_size=(64, 64), energy_ev=800.,
                                  psize_cm=0.67e-7, theta=0.9, n_theta=3, i_theta=1, pos=[(0, 0)], n_modes=1,
                                  free_prop_cm=2e-5, unknown_type='delta_beta', delta_scale=2e-3, beta_scale=3e-4, seed=36,
                                  loss='lsq', raw_data_type='intensity', optimize_prj_pos_offset=True,
                                  prj_pos_offset=(0.7, -0.4)),
    # binning = 3 over 50 slices (propagate.py:207-254: 17 propagation steps, the last bin holds two slices), rotation
    'ptycho3d_p72_bin3': dict(kind='ptycho', obj_size=(48, 50, 50), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                              theta=0.9, n_theta=4, i_theta=2, pos=[(-12, -12), (-30, 5), (1, -19)], n_modes=1,
                              free_prop_cm='inf', unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=26,
                              loss='lsq', raw_data_type='magnitude', binning=3),
    # pure projection (propagate.py:158-186): one modulation by the line integrals, no in-object propagation
    'ptycho3d_p72_pureproj': dict(kind='ptycho', obj_size=(48, 48, 48), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                                  theta=-0.6, n_theta=4, i_theta=1, pos=[(-12, -12), (-30, 5), (1, -19)], n_modes=2,
                                  free_prop_cm='inf', unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=27,
                                  loss='lsq', raw_data_type='magnitude', optimize_probe=True, pure_projection=True),
    # sparse multislice (SparseMultisliceModel, forward_model.py:589-806): three thin slices at unequal positions, one
    # paraxial transfer function per gap; the object is handed over in the beam frame (rotate_out_of_loop)
    'sparse3_p72': dict(kind='sparse', obj_size=(90, 84, 3), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                        theta=0.0, n_theta=1, i_theta=0, pos=[(-10, -10), (20, 13), (9, -12), (0, 4)], n_modes=2,
                        free_prop_cm='inf', unknown_type='delta_beta', delta_scale=3e-2, beta_scale=3e-3, seed=23,
                        loss='lsq', raw_data_type='magnitude', optimize_probe=True, optimize_all_probe_pos=True,
                        slice_pos_cm=[0., 2.0e-6, 7.5e-6], rotate_out_of_loop=True),
    # tilt refinement (optimize_tilt, forward_model.py:274-276: three differentiable rotations instead of the lookup
    # table) + per-angle probe position offset (optimize_probe_pos_offset, forward_model.py:258-260)
    'ptycho3d_p72_tilt': dict(kind='ptycho', obj_size=(48, 48, 48), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                              theta=0.7, n_theta=4, i_theta=1, pos=[(-12, -12), (-30, 5), (1, -19)], n_modes=1,
                              free_prop_cm='inf', unknown_type='delta_beta', delta_scale=8e-4, beta_scale=8e-5, seed=28,
                              loss='lsq', raw_data_type='magnitude', optimize_tilt=True, tilt=(0.7, 0.06, -0.05),
                              optimize_probe_pos_offset=True, probe_pos_offset=(0.4, -0.3)),
    # probes the shared-memory FFT cannot hold (128 pixels): the batched large-frame slice steps.  2-D ptychography with two
    # modes, probe + position refinement, patches overhanging the object on every side, far field
    'ptycho2d_p128': dict(kind='ptycho', obj_size=(160, 150, 1), probe_size=(128, 128), energy_ev=5000., psize_cm=1e-7,
                          theta=0.0, n_theta=1, i_theta=0, pos=[(-20, -30), (40, 31), (5, -3)], n_modes=2,
                          free_prop_cm='inf', unknown_type='delta_beta', delta_scale=3e-2, beta_scale=3e-3, seed=29,
                          loss='lsq', raw_data_type='magnitude', optimize_probe=True, optimize_all_probe_pos=True),
    # 3-D (rotation, 5 slices), 128-pixel probe, near-field detection, intensity data
    'ptycho3d_p128_near': dict(kind='ptycho', obj_size=(140, 132, 5), probe_size=(128, 128), energy_ev=5000., psize_cm=1e-7,
                               theta=0.6, n_theta=3, i_theta=1, pos=[(-10, 9), (20, -14)], n_modes=1,
                               free_prop_cm=2e-6, unknown_type='delta_beta', delta_scale=3e-3, beta_scale=3e-4, seed=30,
                               loss='lsq', raw_data_type='intensity', optimize_probe=True, probe_background=0.4),
    # sparse multislice with REFINABLE slice positions (optimize_slice_pos: the per-gap paraxial kernels stay in the graph,
    # propagate.py:479-534, :556-568)
    'sparse3_p72_slicepos': dict(kind='sparse', obj_size=(90, 84, 3), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                                 theta=0.0, n_theta=1, i_theta=0, pos=[(-10, -10), (20, 13), (9, -12)], n_modes=1,
                                 free_prop_cm='inf', unknown_type='delta_beta', delta_scale=3e-2, beta_scale=3e-3, seed=37,
                                 loss='lsq', raw_data_type='magnitude', optimize_probe=True,
                                 slice_pos_cm=[0., 2.0e-6, 7.5e-6], rotate_out_of_loop=True, optimize_slice_pos=True),
    # 2-D ptychography with probe modes + probe and position refinement (BASELINE config B shape, small)
    'ptycho2d_modes': dict(kind='ptycho', obj_size=(128, 128, 1), probe_size=(72, 72), energy_ev=5000., psize_cm=1e-7,
                           theta=0.0, n_theta=1, i_theta=0,
                           pos=[(-10, -10), (20, 31), (56, 56), (60, -12), (0, 40), (33, 7)], n_modes=3,
                           free_prop_cm='inf', unknown_type='delta_beta', delta_scale=3e-2, beta_scale=3e-3, seed=13,
                           loss='lsq', raw_data_type='magnitude', optimize_probe=True, optimize_all_probe_pos=True),
    # 2-D ptychography, real_imag unknown, P=64, intensity data (BASELINE config C shape, small)
    'ptycho2d_realimag_p64': dict(kind='ptycho', obj_size=(100, 90, 1), probe_size=(64, 64), energy_ev=8801.12,
                                  psize_cm=1.3279e-6, theta=0.0, n_theta=1, i_theta=0,
                                  pos=[(-8, -8), (10, 12), (36, 26), (40, -5), (3, 30)], n_modes=2,
                                  free_prop_cm='inf', unknown_type='real_imag', delta_scale=0.3, beta_scale=0.3, seed=14,
                                  loss='lsq', raw_data_type='intensity', optimize_probe=True,
                                  optimize_all_probe_pos=True),
    # full-field multislice tomography, the reference CI config's shape (BASELINE config A)
    'fullfield64': dict(kind='fullfield', obj_size=(64, 64, 64), probe_size=(64, 64), energy_ev=800., psize_cm=0.67e-7,
                        theta=2.3, n_theta=5, i_theta=2, pos=[(0, 0)], n_modes=1, free_prop_cm=0,
                        unknown_type='delta_beta', delta_scale=2e-3, beta_scale=3e-4, seed=15,
                        loss='lsq', raw_data_type='magnitude', alpha_d=1e-9 * 64 ** 3, alpha_b=1e-10 * 64 ** 3,
                        gamma=1e-6),
}


def make_inputs(name):
    c = dict(CASES[name])
    rs = np.random.RandomState(c['seed'])
    Y, X, Z = c['obj_size']
    P = c['probe_size']
    M = c['n_modes']
    if c['unknown_type'] == 'delta_beta':
        d = _smooth_noise(rs, (Y, X, Z), c['delta_scale'])
        b = _smooth_noise(rs, (Y, X, Z), c['beta_scale'])
        obj = np.stack([d, b], -1).astype('float32')
    else:
        mag = 1.0 - _smooth_noise(rs, (Y, X, Z), c['delta_scale'])
        ph = _smooth_noise(rs, (Y, X, Z), c['beta_scale'] * 3)
        obj = np.stack([mag * np.cos(ph), mag * np.sin(ph)], -1).astype('float32')
    if c['kind'] == 'fullfield':
        pr = np.ones((1,) + tuple(P), 'float32')
        pi = np.zeros((1,) + tuple(P), 'float32')
    else:
        r0, i0 = _gaussian_probe(P, 6.0, 0.5, 6.0)
        pr = np.stack([r0 * (1 + 0.3 * m) + 0.05 * rs.normal(size=P) * np.abs(r0) for m in range(M)]).astype('float32')
        pi = np.stack([i0 * (1 - 0.2 * m) + 0.05 * rs.normal(size=P) * np.abs(r0) for m in range(M)]).astype('float32')
        if c.get('probe_background'):
            # a plane-wave pedestal: without it the near-field magnitude is ~0 away from the beam and the phase of a
            # vanishing field -- hence the gradient there -- is rounding noise in fp32 (reference included)
            pr = (pr + np.float32(c['probe_background'])).astype('float32')
    pos = np.array(c['pos'], dtype='int64')
    B = len(pos)
    # synthetic "measured" data: positive, smooth-ish, same order of magnitude as a plausible prediction
    amp = float(np.sqrt((pr ** 2 + pi ** 2).sum()))
    if c['free_prop_cm'] == 'inf' and not c.get('normalize_fft', False) and c.get('sign_convention', 1) == 1:
        scale = amp * 2.0
    else:
        scale = amp / np.sqrt(P[0] * P[1]) * 2.0
    prj = (scale * rs.uniform(0.2, 1.0, size=(B,) + tuple(P))).astype('float32')
    if c['loss'] == 'poisson':
        # Poisson loss has a 1/a term: keep the case well conditioned by giving the measured data the
        # far-field envelope of the probe behind a random phase screen (instead of white noise).
        screen = np.exp(1j * _smooth_noise(rs, tuple(P), 2.0))
        ff = np.fft.fftshift(np.fft.ifft2((pr[0] + 1j * pi[0]) * screen, norm='ortho'
                                          if c.get('normalize_fft', False) else None))
        env = np.abs(ff) + 1e-3 * np.abs(ff).max()
        prj = (env[None] * rs.uniform(0.7, 1.3, size=(B,) + tuple(P))).astype('float32')
    if c['raw_data_type'] == 'intensity':
        prj = (prj ** 2).astype('float32')
    n_theta = c['n_theta']
    ppc = None
    if c.get('optimize_all_probe_pos', False):
        ppc = rs.uniform(-0.6, 0.6, size=(n_theta, 16, 2)).astype('float32')
    ind = np.arange(B) + 2 if ppc is not None else np.arange(B)
    out = dict(cfg=c, obj=obj, probe_real=pr, probe_imag=pi, pos=pos, prj=prj, probe_pos_correction=ppc,
               this_ind_batch=ind, weight_l1=None, beamstop=None)
    if c.get('reweighted', False):
        # ptychography.py:995-999: weight_l1 = max(obj) / (|obj| + 1e-4 mean(obj)), in fp32 as the driver does
        out['weight_l1'] = (obj.max() / (np.abs(obj) + np.float32(1e-4) * obj.mean())).astype('float32')
    if c.get('beamstop_radius'):
        yy, xx = np.meshgrid(np.arange(P[0]) - P[0] // 2, np.arange(P[1]) - P[1] // 2, indexing='ij')
        out['beamstop'] = (np.sqrt(yy ** 2 + xx ** 2) > c['beamstop_radius']).astype('float32')   # 1 = valid pixel
    return out


# ---------------------------------------------------------------------------------------------------------
# Multi-distance holography (MultiDistModel, BASELINE config E shape, small frames): one thin object, D distances
# ---------------------------------------------------------------------------------------------------------
MD_CASES = {
    # demos/2d_multidist_holography_w_position_correction.py: real_imag object, plane probe, intensity data, lsq,
    # refinable per-distance shift of the measured frames
    'multidist256_shift': dict(N=256, energy_ev=17500., psize_cm=1e-4, dists=[40., 60., 80., 100.], n_modes=1,
                               unknown_type='real_imag', delta_scale=0.4, beta_scale=0.3, seed=31, loss='lsq',
                               raw_data_type='intensity', shift=True, optimize_probe=False),
    # delta_beta object, structured probe, magnitude data, probe gradient (the reference's MultiDistModel indexes the
    # imaginary probe as [i_mode] on the batch axis, forward_model.py:1005, so it only runs with one mode -- quirk 9)
    'multidist128_probe': dict(N=128, energy_ev=17500., psize_cm=1e-4, dists=[30., 70.], n_modes=1,
                               unknown_type='delta_beta', delta_scale=2e-7, beta_scale=2e-8, seed=32, loss='lsq',
                               raw_data_type='magnitude', shift=False, optimize_probe=True),
    # refinable propagation distances (optimize_free_prop: in-graph transfer function, propagate.py:84-103,271-276)
    'multidist128_dist': dict(N=128, energy_ev=17500., psize_cm=1e-4, dists=[35., 62.], n_modes=1,
                              unknown_type='real_imag', delta_scale=0.4, beta_scale=0.3, seed=33, loss='lsq',
                              raw_data_type='intensity', shift=False, optimize_probe=False, free_prop=True),
    # refinable affine alignment of the measured holograms (optimize_prj_affine, forward_model.py:1066-1072;
    # demos/2d_multidist_holography_w_affine.py)
    'multidist128_affine': dict(N=128, energy_ev=17500., psize_cm=1e-4, dists=[35., 62., 90.], n_modes=1,
                                unknown_type='real_imag', delta_scale=0.4, beta_scale=0.3, seed=34, loss='lsq',
                                raw_data_type='intensity', shift=False, optimize_probe=False, affine=True),
    # frames sub-divided into 3 x 3 tiles of 48 pixels, propagated with a safe zone of 8 pixels (64-pixel wavefields), a
    # minibatch of 4 tiles in sub-batches of 2 (forward_model.py:917-1026); structured probe so that the reference's
    # imaginary-patch indexing (quirk 9) is visible
    'multidist144_tiled': dict(N=144, energy_ev=17500., psize_cm=1e-4, dists=[30., 70.], n_modes=1,
                               unknown_type='real_imag', delta_scale=0.4, beta_scale=0.3, seed=35, loss='lsq',
                               raw_data_type='intensity', shift=False, optimize_probe=True,
                               tiles=dict(sub=48, szw=8, ind=[1, 3, 4, 8], n_dp_batch=2)),
}


def md_tile_positions(N, sub):
    n = N // sub
    return np.array([(sub * (i // n), sub * (i % n)) for i in range(n * n)], dtype='int64')


def make_md_inputs(name):
    c = dict(MD_CASES[name])
    rs = np.random.RandomState(c['seed'])
    N, M, D = c['N'], c['n_modes'], len(c['dists'])
    if c['unknown_type'] == 'delta_beta':
        d = _smooth_noise(rs, (N, N, 1), c['delta_scale'])
        b = _smooth_noise(rs, (N, N, 1), c['beta_scale'])
        obj = np.stack([d, b], -1).astype('float32')
    else:
        mag = 1.0 - _smooth_noise(rs, (N, N, 1), c['delta_scale'])
        ph = _smooth_noise(rs, (N, N, 1), c['beta_scale'] * 3)
        obj = np.stack([mag * np.cos(ph), mag * np.sin(ph)], -1).astype('float32')
    if not c['optimize_probe']:
        pr = np.ones((1, N, N), 'float32')          # probe_type='plane'
        pi = np.zeros((1, N, N), 'float32')
    else:
        pr = np.stack([1.0 / (1 + m) + 0.2 * _smooth_noise(rs, (N, N), 1.0) for m in range(M)]).astype('float32')
        pi = np.stack([0.3 * _smooth_noise(rs, (N, N), 1.0) - 0.1 * m for m in range(M)]).astype('float32')
    amp = float(np.sqrt((pr ** 2 + pi ** 2).sum(0).mean()))
    prj = (amp * (0.5 + 0.5 * np.stack([_smooth_noise(rs, (N, N), 1.0) for _ in range(D)]))).astype('float32')
    if c['raw_data_type'] == 'intensity':
        prj = (prj ** 2).astype('float32')
    if c.get('tiles'):
        # exchange/data of a tiled acquisition: [n_dists * n_blocks, sub, sub], distance-major, blocks in raster order
        sub = c['tiles']['sub']
        tp = md_tile_positions(N, sub)
        prj = np.stack([prj[d, y:y + sub, x:x + sub] for d in range(D) for (y, x) in tp]).astype('float32')
    ppc = rs.uniform(-0.8, 0.8, size=(D, 2)).astype('float32') if c['shift'] else None
    aff = None
    if c.get('affine', False):
        ident = np.array([[1., 0, 0], [0, 1., 0]])
        aff = (ident[None] + rs.uniform(-0.03, 0.03, size=(D, 2, 3))).astype('float32')
        aff[0] = ident               # the driver pins the first frame's transformation (optimizers.py:1069-1074)
    return dict(cfg=c, obj=obj, probe_real=pr, probe_imag=pi, prj=prj, probe_pos_correction=ppc, prj_affine_ls=aff)
