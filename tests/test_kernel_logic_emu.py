"""CPU: the kernel SOURCES (adorym_b200/csrc/*.cu) compiled as C++ under a host emulation of the CUDA execution
model (tests/emu/) and checked against the oracle / the reference goldens.  This validates index permutations,
butterflies, adjoints and the C ABI in the GPU-less container; the real parity tests are tests/test_gpu_parity.py."""
import os
import sys
import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
from oracle_util import case_cfg, run_oracle_case, rel_err, compact_grad  # noqa: E402
from oracle import multislice_oracle as orc  # noqa: E402
import emu_lib  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def run_emu_case(name, separable=False, **kw):
    inp, cfg, coord, reg = case_cfg(name)
    c = inp['cfg']
    Y, X, Z = c['obj_size']
    grid = None
    if coord is not None:
        grid = np.ascontiguousarray(orc.grid_normalise(coord, X, Z).numpy(), dtype='float32')
    voxel_nm = c['psize_cm'] * 1e7
    k1 = 2. * orc.PI * voxel_nm / (1240. / c['energy_ev'])
    shift = None
    if inp['probe_pos_correction'] is not None:
        shift = inp['probe_pos_correction'][c['i_theta'], inp['this_ind_batch']]
    fp = 'inf' if c['free_prop_cm'] == 'inf' else ('none' if c['free_prop_cm'] in (0, None) else 'near')
    if c.get('pure_projection', False):
        kw = dict(kw, binning=Z)
    elif c.get('binning', 1) > 1:
        kw = dict(kw, binning=c['binning'])
    if c.get('slice_pos_cm') is not None:
        gaps = np.diff(np.array(c['slice_pos_cm'])) * 1e7
        kw = dict(kw, h_gaps=[orc.get_kernel(float(g), 1240. / c['energy_ev'], np.array([voxel_nm] * 3), c['probe_size'])
                              for g in gaps])
    return inp, emu_lib.loss_and_grads_emu(
        inp['obj'], inp['probe_real'], inp['probe_imag'], inp['pos'], inp['prj'], grid, cfg['h'], k1,
        unknown_type=c['unknown_type'], free_prop=fp, sign_convention=c.get('sign_convention', 1),
        normalize_fft=c.get('normalize_fft', False), loss=c['loss'], raw_data_type=c['raw_data_type'],
        poisson_multiplier=c.get('poisson_multiplier', 1.), shift=shift, reg=reg, separable=separable,
        weight_l1=inp.get('weight_l1'), beamstop=inp.get('beamstop'), **kw)


@pytest.mark.parametrize('name', ['ptycho2d_modes', 'ptycho2d_realimag_p64', 'fullfield64', 'ptycho3d_p72',
                                  'ptycho3d_p72_poisson', 'ptycho3d_p72_edge', 'ptycho3d_p72_exitwave',
                                  'ptycho3d_p72_strongreg', 'ptycho3d_p72_rwl1', 'ptycho3d_p72_beamstop',
                                  'sparse3_p72', 'ptycho3d_p72_bin3', 'ptycho3d_p72_pureproj'])
def test_emulated_kernels_match_oracle_and_golden(name):
    # the separable-H path (Fresnel kernel in two shared-memory tables) on the 3-D P=72 cases, the general-H path elsewhere
    # (fullfield64: the Cooley-Tukey 64-point transform on the separable five-pass path)
    inp, out = run_emu_case(name, separable=name.startswith('ptycho3d') or name == 'fullfield64')
    _, loss, pred, grads = run_oracle_case(name, torch.float64)
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    gtol = 1e-3 if 'poisson' in name else 1e-4
    assert abs(out['loss'] - loss) <= 1e-4 * abs(loss)
    assert abs(out['loss'] - float(gold['loss'])) <= 1e-4 * abs(float(gold['loss']))
    assert rel_err(out['pred'], pred) < 1e-4
    assert rel_err(out['pred'], gold['pred']) < 1e-4
    assert rel_err(out['g_obj'], grads['obj']) < gtol
    cg = compact_grad(out['g_obj'])
    for k in ('proj0', 'proj1', 'proj2', 'sample'):
        assert rel_err(cg[k], gold['gobj_' + k]) < gtol, k
    if 'probe_real' in grads:
        assert rel_err(out['g_probe_real'], grads['probe_real']) < 1e-4
        assert rel_err(out['g_probe_imag'], grads['probe_imag']) < 1e-4
        assert rel_err(out['g_probe_real'], gold['g_probe_real']) < 1e-4
    if 'probe_pos_correction' in grads:
        c = inp['cfg']
        g_ref = grads['probe_pos_correction'][c['i_theta'], inp['this_ind_batch']]
        assert rel_err(out['g_shift'], g_ref) < 1e-4


@pytest.mark.parametrize('name,grid', [('ptycho3d_p72', 3), ('ptycho3d_p72_edge', 3), ('ptycho3d_p72_edge', 7),
                                       ('ptycho3d_p72_exitwave', 2), ('ptycho2d_realimag_p64', 4)])
def test_emulated_balanced_schedule_matches_one_cta_per_wavefield(name, grid, monkeypatch):
    """The persistent balanced schedule (multislice.cu: equal intervals of slice steps per CTA, wavefields cut by an
    interval boundary migrate between CTAs through the workspace) must give the results of the one-CTA-per-wavefield
    launch: ADM_MS_GRID forces fewer co-resident CTAs than wavefields, so chains are cut at arbitrary slices
    (the 2-D case has one slice per wavefield: whole wavefields per CTA, no cut)."""
    inp, ref = run_emu_case(name, separable=name.startswith('ptycho3d'))
    monkeypatch.setenv('ADM_MS_GRID', str(grid))
    _, out = run_emu_case(name, separable=name.startswith('ptycho3d'))
    for k in ('det', 'pred', 'g_obj', 'g_probe_real', 'g_probe_imag'):
        assert rel_err(out[k], ref[k]) < 2e-6, k      # the TMA reduce-adds arrive in another order: fp32 sum order only
    assert abs(out['loss'] - ref['loss']) <= 1e-6 * abs(ref['loss'])


@pytest.mark.parametrize('name', ['ptycho3d_p72_edge', 'sparse3_p72', 'fullfield64'])
def test_emulated_five_pass_kernels_match_three_pass_kernels(name, monkeypatch):
    """Probes of 72 / 64 pixels with a separable transfer function run on the three-pass kernels (ms3_kernels.cuh:
    72 = 3 x 24, 64 = 4 x 16) by default, and every such case above exercises them; ADM_MS3=0 selects the five-pass
    kernels of multislice.cu (72 = 8 x 9, 64 = 8 x 8), which stay the path for near-field detection / general H /
    adjoint-wavefield checkpoints.  Two independent factorizations of the same transform must agree to fp32 rounding."""
    _, a = run_emu_case(name, separable=True)
    monkeypatch.setenv('ADM_MS3', '0')
    _, b = run_emu_case(name, separable=True)
    for k in ('det', 'pred', 'g_obj', 'g_probe_real', 'g_probe_imag'):
        assert 0 < rel_err(a[k], b[k]) < 5e-5, k      # 0 would mean that the switch selected nothing
    assert abs(a['loss'] - b['loss']) <= 1e-5 * abs(b['loss'])
    if name.startswith('ptycho3d_p72'):
        # the forward kernel has two thread configurations: 7 warps (few wavefields, as here) and 9 warps with the M pass in
        # two rounds and the y / x twiddles combined into one factor per element (more wavefields than SMs); ADM_MS3_FWD=2
        # forces the second one: same blocks, another thread mapping, another rounding of the twiddles
        monkeypatch.setenv('ADM_MS3', '1')
        monkeypatch.setenv('ADM_MS3_FWD', '2')
        _, c = run_emu_case(name, separable=True)
        for k in ('det', 'pred', 'g_obj'):
            assert 0 < rel_err(c[k], a[k]) < 5e-5, k


def test_emulated_three_pass_kernels_for_64_pixel_probes_with_handoff(monkeypatch):
    """64-pixel probes on the three-pass kernels (64 = 4 x 16: 4 x 4 blocks in the M pass, 16-point transforms in the B
    passes): several overhanging wavefields through 5 slices, far-field detection, probe gradient; (a) against the five-pass
    kernels, (b) with the chains cut at arbitrary slices and handed from CTA to CTA (ADM_MS_GRID)."""
    rs = np.random.RandomState(11)
    Y, X, Z, P, B = 90, 83, 5, 64, 4
    obj = np.zeros((Y, X, Z, 2), 'float32')
    obj[..., 0] = 3e-3 * rs.uniform(size=(Y, X, Z))
    obj[..., 1] = 3e-4 * rs.uniform(size=(Y, X, Z))
    yy, xx = np.mgrid[:P, :P] - P / 2
    amp = np.exp(-(yy ** 2 + xx ** 2) / (2 * 12. ** 2))
    ph = 0.3 * np.sin(xx / 7.) + 0.2 * np.cos(yy / 5.)
    pr = (amp * np.cos(ph))[None].astype('float32')
    pi = (amp * np.sin(ph))[None].astype('float32')
    pos = np.array([[-9, 30], [20, -5], [31, 27], [5, 3]], 'int32')
    prj = rs.uniform(0.5, 1.5, size=(B, P, P)).astype('float32')
    voxel, lam = 1.0, 1240. / 5000
    h = orc.get_kernel(voxel, lam, np.array([voxel] * 3), [P, P], fresnel_approx=True, sign_convention=1)
    k1 = 2. * orc.PI * voxel / lam
    kw = dict(unknown_type='delta_beta', free_prop='inf', loss='lsq', raw_data_type='magnitude', separable=True)
    a = emu_lib.loss_and_grads_emu(obj, pr, pi, pos, prj, None, h, k1, **kw)
    monkeypatch.setenv('ADM_MS_GRID', '3')
    c = emu_lib.loss_and_grads_emu(obj, pr, pi, pos, prj, None, h, k1, **kw)
    monkeypatch.delenv('ADM_MS_GRID')
    monkeypatch.setenv('ADM_MS3', '0')
    b = emu_lib.loss_and_grads_emu(obj, pr, pi, pos, prj, None, h, k1, **kw)
    for k in ('det', 'pred', 'g_obj', 'g_probe_real', 'g_probe_imag'):
        assert 0 < rel_err(a[k], b[k]) < 5e-5, k
        assert rel_err(c[k], a[k]) < 2e-6, k
    assert abs(a['loss'] - b['loss']) <= 1e-5 * abs(b['loss'])


@pytest.mark.parametrize('P', [72, 64])
def test_emulated_three_pass_kernels_near_field_detection(P, monkeypatch):
    """Near-field detection (finite free_prop_cm, propagate.py:271-280) on the three-pass kernels: the free-space transfer
    function arrives as the general prepared table and, being an outer product (cfg.hfree_separable), is applied as two more
    B passes with the factors taken from its first row / column; forward, object and probe gradients against the five-pass
    kernels (which are pinned by the oracle in tests/test_gpu_parity.py::test_near_field_gradients_match_oracle)."""
    rs = np.random.RandomState(13)
    Y, X, Z, B = P + 21, P + 9, 4, 3
    obj = np.zeros((Y, X, Z, 2), 'float32')
    obj[..., 0] = 3e-3 * rs.uniform(size=(Y, X, Z))
    obj[..., 1] = 3e-4 * rs.uniform(size=(Y, X, Z))
    pr = rs.normal(size=(1, P, P)).astype('float32')
    pi = rs.normal(size=(1, P, P)).astype('float32')
    pos = np.array([[-7, 12], [15, -4], [9, 6]], 'int32')
    prj = rs.uniform(0.5, 1.5, size=(B, P, P)).astype('float32')
    voxel, lam = 1.0, 1240. / 5000
    h = orc.get_kernel(voxel, lam, np.array([voxel] * 3), [P, P], fresnel_approx=True, sign_convention=1)
    hf = orc.get_kernel(2e-4 * 1e7, lam, np.array([voxel] * 3), [P, P], fresnel_approx=True, sign_convention=1)
    k1 = 2. * orc.PI * voxel / lam
    kw = dict(unknown_type='delta_beta', free_prop='near', Hfree=hf, loss='lsq', raw_data_type='magnitude', separable=True)
    a = emu_lib.loss_and_grads_emu(obj, pr, pi, pos, prj, None, h, k1, **kw)
    monkeypatch.setenv('ADM_MS3', '0')
    b = emu_lib.loss_and_grads_emu(obj, pr, pi, pos, prj, None, h, k1, **kw)
    for k in ('det', 'pred', 'g_obj', 'g_probe_real', 'g_probe_imag'):
        assert 0 < rel_err(a[k], b[k]) < 5e-5, k
    assert abs(a['loss'] - b['loss']) <= 1e-5 * abs(b['loss'])


@pytest.mark.parametrize('name', ['ptycho3d_p72_strongreg', 'fullfield64', 'ptycho2d_realimag_p64'])
def test_emulated_fused_object_chain_matches_separate_kernels(name):
    """object_chain.cu (rotation + transmission in one pass, regulariser gradient as the initial value of the gradient
    buffer, y slabs) against the separate rotate.cu / transmission.cu / adam.cu kernels: same arithmetic, so the results
    agree to fp32 summation order."""
    sep = name.startswith('ptycho3d')
    _, ref = run_emu_case(name, separable=sep, fused_chain=False)
    _, out = run_emu_case(name, separable=sep, fused_chain=True, grad_slabs=3)
    assert abs(out['loss'] - ref['loss']) <= 1e-6 * abs(ref['loss'])
    for k in ('pred', 'g_obj', 'g_probe_real'):
        assert rel_err(out[k], ref[k]) < 2e-6, k


@pytest.mark.parametrize('name', ['multidist128_probe', 'multidist256_shift'])
def test_emulated_large_frame_kernels_match_oracle_and_golden(name):
    """MultiDistModel path: large-frame FFT passes (digit-reversed spectrum, H permutation, adjoint summed over
    distances), measured-frame shift and its gradient, large detector loss -- emulated kernels vs the fp64 oracle and
    the unmodified reference's golden."""
    from oracle_util import run_md_oracle_case, check_md_against_golden, md_case_cfg
    inp, _ = md_case_cfg(name)
    c = inp['cfg']
    N = c['N']
    voxel_nm = np.array([c['psize_cm']] * 3) * 1e7
    lmbda_nm = 1240. / c['energy_ev']
    hs = [orc.get_kernel(d * 1e7, lmbda_nm, voxel_nm, [N, N]) for d in c['dists']]
    k1 = 2. * orc.PI * voxel_nm[-1] / lmbda_nm
    out = emu_lib.multidist_loss_and_grads_emu(inp['obj'], inp['probe_real'], inp['probe_imag'], inp['prj'], hs, k1,
                                               unknown_type=c['unknown_type'], loss=c['loss'],
                                               raw_data_type=c['raw_data_type'], shift=inp['probe_pos_correction'])
    _, loss, pred, grads = run_md_oracle_case(name, torch.float64)
    assert abs(out['loss'] - loss) <= 1e-4 * abs(loss)
    assert rel_err(out['pred'], pred) < 1e-4
    assert rel_err(out['g_obj'], grads['obj']) < 1e-4
    mine = {'obj': out['g_obj'], 'probe_real': out['g_probe_real'], 'probe_imag': out['g_probe_imag']}
    if 'probe_real' in grads:
        assert rel_err(out['g_probe_real'], grads['probe_real']) < 1e-4
        assert rel_err(out['g_probe_imag'], grads['probe_imag']) < 1e-4
    if 'probe_pos_correction' in grads:
        assert rel_err(out['g_shift'], grads['probe_pos_correction']) < 1e-4
        mine['probe_pos_correction'] = out['g_shift']
    gold = np.load(os.path.join(GOLD, name + '.npz'))
    check_md_against_golden(gold, out['loss'], out['pred'], mine)


def test_emulated_adam_matches_reference_golden_bitwise():
    """adm_adam_step follows optimizers.py:309-318 op by op in fp32 (no FMA contraction, true divisions): the emulated
    kernel must reproduce the reference's three-call golden to the last bit (m, v) / 1 ulp (x)."""
    import ctypes
    from make_golden import adam_inputs
    L = emu_lib.emu()
    gold = np.load(os.path.join(GOLD, 'adam.npz'))
    x, g_ls, steps = adam_inputs()
    x = np.ascontiguousarray(x, dtype='float32')
    m = np.zeros_like(x)
    v = np.zeros_like(x)
    for i, (g, st) in enumerate(zip(g_ls, steps)):
        g = np.ascontiguousarray(g, dtype='float32')
        rc = L.adm_adam_step(emu_lib.P(x), emu_lib.P(m), emu_lib.P(v), emu_lib.P(g), x.size, 1e-5, 0.9, 0.999, 1e-7, st, 0,
                             None, None)
        assert rc == 0
        assert np.array_equal(m, gold['m%d' % i])
        assert np.array_equal(v, gold['v%d' % i])
        np.testing.assert_allclose(x, gold['x%d' % i], rtol=2e-7, atol=0)


def test_emulated_peer_memory_update_equals_allreduce_plus_adam():
    """adm_p2p_reduce_adam (the data-parallel object update over NVLink peer memory): three emulated ranks, each pulling
    its shard of every rank's gradient, applying Adam and pushing the shard into every rank's object, must leave every
    copy of the object equal -- bit for bit -- to rank-ordered sum + adm_adam_step on the whole volume."""
    import ctypes
    L = emu_lib.emu()
    P = emu_lib.P
    rs = np.random.RandomState(8)
    world, n = 3, 4 * 1234
    x0 = (rs.normal(size=n) * 1e-3).astype('float32')
    gs = [(rs.normal(size=n) * 10.0 ** rs.uniform(-6, 0, size=n)).astype('float32') for _ in range(world)]
    mask = (rs.uniform(size=n // 2) > 0.3).astype('float32')
    # reference: sum in rank order, one Adam step with the epilogue (clip + mask)
    gsum = gs[0].copy()
    for g in gs[1:]:
        gsum = gsum + g
    xr, mr, vr = x0.copy(), np.zeros(n, 'float32'), np.zeros(n, 'float32')
    assert L.adm_adam_step(P(xr), P(mr), P(vr), P(gsum), n, 1e-3, 0.9, 0.999, 1e-7, 2, 1, P(mask), None) == 0
    xs = [x0.copy() for _ in range(world)]
    ms = [np.zeros(n, 'float32') for _ in range(world)]
    vs = [np.zeros(n, 'float32') for _ in range(world)]
    xp = (ctypes.c_void_p * world)(*[a.ctypes.data for a in xs])
    gp = (ctypes.c_void_p * world)(*[a.ctypes.data for a in gs])
    for r in range(world):
        rc = L.adm_p2p_reduce_adam(ctypes.cast(xp, ctypes.c_void_p), ctypes.cast(gp, ctypes.c_void_p), None, None, world, r,
                                   P(ms[r]), P(vs[r]), n, 1e-3, 0.9, 0.999, 1e-7, 2, 1, P(mask), None)
        assert rc == 0
    for r in range(world):
        assert np.array_equal(xs[r], xr), r
        i0, i1 = 4 * ((n // 4) * r // world), 4 * ((n // 4) * (r + 1) // world)
        assert np.array_equal(ms[r][i0:i1], mr[i0:i1]) and np.array_equal(vs[r][i0:i1], vr[i0:i1])


@pytest.mark.parametrize('shape,C,roles', [((3, 16, 20), 1, (0, 1, 2)), ((6, 10, 12), 2, (0, 2, 1)),
                                            ((6, 10, 12), 2, (1, 2, 0)), ((6, 10, 12), 2, (2, 0, 1))])
def test_emulated_affine_resample_matches_oracle(shape, C, roles):
    """affine.cu (tilt rotation / affine alignment of the data: wrappers.py:1158-1217) against the oracle's restatement
    of affine_grid + grid_sample: value, transpose and matrix gradient, per-plane and shared matrices, all three axis
    roles of wrappers.rotate."""
    import ctypes
    L = emu_lib.emu()
    P = emu_lib.P
    rs = np.random.RandomState(5)
    full = shape + ((2,) if C == 2 else ())
    arr = rs.normal(size=full).astype('float32')
    w = rs.normal(size=full).astype('float32')
    N = shape[roles[0]]
    perm = list(roles) + ([3] if C == 2 else [])
    inv = [perm.index(i) for i in range(len(perm))]
    for per_plane in (1, 0):
        nm = N if per_plane else 1
        mats = (np.array([[1., 0, 0], [0, 1., 0]]) + 0.15 * rs.normal(size=(nm, 2, 3))).astype('float32')
        out = np.full(full, np.nan, 'float32')
        emu_lib.ck(L.adm_affine_resample_fwd(shape[0], shape[1], shape[2], C, roles[0], roles[1], roles[2], P(arr), P(mats),
                                             per_plane, P(out), None))
        g_in = np.zeros(full, 'float32')
        emu_lib.ck(L.adm_affine_resample_adj(shape[0], shape[1], shape[2], C, roles[0], roles[1], roles[2], P(w), P(mats),
                                             per_plane, P(g_in), None))
        g_m = np.zeros((N, 6), 'float32')
        emu_lib.ck(L.adm_affine_resample_dmat(shape[0], shape[1], shape[2], C, roles[0], roles[1], roles[2], P(arr), P(w),
                                              P(mats), per_plane, P(g_m), None))
        a = torch.tensor(arr, dtype=torch.float64, requires_grad=True)
        m = torch.tensor(mats, dtype=torch.float64, requires_grad=True)
        planes = a.permute(*perm) if C == 2 else a.permute(*perm).unsqueeze(-1)
        ref = orc.affine_resample(planes, m)
        ref = ref.permute(*inv) if C == 2 else ref[..., 0].permute(*inv)
        ga, gm = torch.autograd.grad((ref * torch.tensor(w, dtype=torch.float64)).sum(), [a, m])
        assert rel_err(out, ref.detach().numpy()) < 1e-5
        assert rel_err(g_in, ga.numpy()) < 1e-5
        got = g_m.reshape(N, 2, 3) if per_plane else g_m.sum(0).reshape(1, 2, 3)
        assert rel_err(got, gm.numpy()) < 1e-4


def test_emulated_large_probe_patch_step_and_far_field():
    """large_frame.cu patch kernels (probes of 128 pixels: adm_lf_patch_step_fwd / _bwd, adm_lf_farfield_fwd / _bwd,
    adm_frame_shift_complex_fwd) against numpy FFTs: value of one slice step for two overhanging patches, its adjoint
    (dL/d psi and the scatter-added dL/dT) by the definition <A x, y> = <x, A^H y>, the far-field transform and its
    adjoint, and the complex Fourier shift (util.py:380-397)."""
    L = emu_lib.emu()
    P = emu_lib.P
    ck = emu_lib.ck
    rs = np.random.RandomState(9)
    N, B, M, Y, X, Nz = 128, 2, 1, 150, 141, 2
    Xp = (X + 1) & ~1
    F = B * M
    tvol = np.zeros((Nz, Y, Xp, 2), 'float32')
    tvol[:, :, :X] = 0.2 * rs.normal(size=(Nz, Y, X, 2))
    pos = np.array([[-20, 30], [40, -9]], 'int32')
    psi = rs.normal(size=(F, N, N, 2)).astype('float32')
    order = np.zeros(N, 'int32')
    ck(L.adm_lf_spectrum_order(N, P(order)))
    H = np.exp(1j * rs.uniform(0, 2 * np.pi, size=(N, N)))
    hp = H[np.ix_(order, order)] / (N * N)
    Hp = np.ascontiguousarray(np.stack([hp.real, hp.imag], -1), dtype='float32')
    Hc = np.ascontiguousarray(Hp * np.array([1, -1], 'float32'))
    z = 1
    c = lambda a: a[..., 0].astype('float64') + 1j * a[..., 1]

    def patch(b):
        t = np.zeros((N, N), 'complex128')
        for yy in range(N):
            gy = pos[b, 0] + yy
            if 0 <= gy < Y:
                x0, x1 = max(0, -pos[b, 1]), min(N, X - pos[b, 1])
                t[yy, x0:x1] = c(tvol[z, gy, pos[b, 1] + x0:pos[b, 1] + x1])
        return t

    work = np.full((F, N, N, 2), np.nan, 'float32')
    out = np.full((F, N, N, 2), np.nan, 'float32')
    ck(L.adm_lf_patch_step_fwd(N, B, M, Y, X, z, P(tvol), P(pos), P(psi), P(Hp), P(work), P(out), None))
    ref = np.stack([np.fft.ifft2(H * np.fft.fft2(c(psi[b]) * (1 + patch(b)))) for b in range(F)])
    assert rel_err(c(out), ref) < 1e-5
    # modulation only (last slice)
    out0 = np.full((F, N, N, 2), np.nan, 'float32')
    ck(L.adm_lf_patch_step_fwd(N, B, M, Y, X, z, P(tvol), P(pos), P(psi), None, None, P(out0), None))
    assert rel_err(c(out0), np.stack([c(psi[b]) * (1 + patch(b)) for b in range(F)])) < 1e-6
    # adjoint: dL/d psi and dL/dT for L = Re <g, out>
    g = rs.normal(size=(F, N, N, 2)).astype('float32')
    w1 = np.full((F, N, N, 2), np.nan, 'float32')
    w2 = np.full((F, N, N, 2), np.nan, 'float32')
    g_psi = np.full((F, N, N, 2), np.nan, 'float32')
    g_tvol = np.zeros((Nz, Y, Xp, 2), 'float32')
    ck(L.adm_lf_patch_step_bwd(N, B, M, Y, X, z, P(tvol), P(pos), P(psi), P(Hc), P(g), P(w1), P(w2), P(g_psi), P(g_tvol), 0, None))
    gm = np.stack([np.fft.ifft2(np.conj(H) * np.fft.fft2(c(g[b]))) for b in range(F)])          # gradient behind the modulation
    ref_gpsi = np.stack([gm[b] * np.conj(1 + patch(b)) for b in range(F)])
    assert rel_err(c(g_psi), ref_gpsi) < 1e-5
    ref_gt = np.zeros((Nz, Y, Xp), 'complex128')
    for b in range(F):
        gt = gm[b] * np.conj(c(psi[b]))
        for yy in range(N):
            gy = pos[b, 0] + yy
            if 0 <= gy < Y:
                x0, x1 = max(0, -pos[b, 1]), min(N, X - pos[b, 1])
                ref_gt[z, gy, pos[b, 1] + x0:pos[b, 1] + x1] += gt[yy, x0:x1]
    assert rel_err(c(g_tvol), ref_gt) < 1e-5
    # far field and its adjoint
    od = np.ascontiguousarray(order)
    ff = np.full((F, N, N, 2), np.nan, 'float32')
    ck(L.adm_lf_farfield_fwd(N, F, P(psi), P(od), 0.5, P(work), P(ff), None))
    assert rel_err(c(ff), 0.5 * np.fft.fftshift(np.fft.fft2(c(psi)), axes=(1, 2))) < 1e-5
    gi = np.full((F, N, N, 2), np.nan, 'float32')
    ck(L.adm_lf_farfield_bwd(N, F, P(g), P(od), 0.5, P(work), P(gi), None))
    lhs = np.vdot(c(g), 0.5 * np.fft.fftshift(np.fft.fft2(c(psi)), axes=(1, 2)))         # <g, A psi>
    rhs = np.vdot(c(gi), c(psi))                                                          # <A^H g, psi>
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs)
    # complex Fourier shift
    sh = np.array([[0.6, -1.3], [-2.2, 0.4]], 'float32')
    spec = np.full((F, N, N, 2), np.nan, 'float32')
    so = np.full((F, N, N, 2), np.nan, 'float32')
    ck(L.adm_frame_shift_complex_fwd(N, F, P(psi), P(sh), P(od), P(work), P(spec), P(so), None))
    fy, fx = np.meshgrid(np.fft.fftfreq(N), np.fft.fftfreq(N), indexing='ij')
    ref_s = np.stack([np.fft.ifft2(np.fft.fft2(c(psi[b])) * np.exp(-2j * orc.PI * (fx * sh[b, 1] + fy * sh[b, 0]))) for b in range(F)])
    assert rel_err(c(so), ref_s) < 1e-5
