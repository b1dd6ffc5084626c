"""TEST INFRASTRUCTURE ONLY: ctypes handle on the host-emulated kernels (tests/emu/libadm_emu.so) and a numpy
orchestration of one loss+gradient evaluation, mirroring what adorym_b200.ops does on the GPU."""
import ctypes
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from build_emu import build  # noqa: E402
from adorym_b200._lib import SIGNATURES, MultisliceCfg, bind  # signature table only  # noqa: E402

_emu = None
PI = 3.14159265359


def emu():
    global _emu
    if _emu is None:
        _emu = bind(ctypes.CDLL(build()))
    return _emu


def P(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ck(rc):
    assert rc == 0, 'emu call failed rc=%d' % rc


def loss_and_grads_emu(obj, probe_real, probe_imag, pos, prj_abs, grid, H, k1, *, unknown_type='delta_beta',
                       free_prop='inf', sign_convention=1, normalize_fft=False, loss='lsq', raw_data_type='magnitude',
                       poisson_multiplier=1.0, shift=None, reg=None, Hfree=None, separable=False, fused_chain=True,
                       grad_slabs=1, weight_l1=None, beamstop=None, h_gaps=None, binning=1):
    """fused_chain: object -> transmission volume and back through the fused kernels (object_chain.cu), as
    ops.MultisliceFn layout 'yxz' does; False: the separate rotate / transmission kernels (rotate.cu, transmission.cu)."""
    L = emu()
    Y, X, Zfull, _ = obj.shape
    Z = -(-Zfull // binning)          # propagation steps
    M, Pp, _ = probe_real.shape
    B = len(pos)
    obj = np.ascontiguousarray(obj, dtype='float32')
    if fused_chain:
        rot = None
    elif grid is not None:
        rot = np.zeros((Z, Y, X, 2), 'float32')
        ck(L.adm_rotate_fwd(Y, X, Z, P(obj), P(grid), P(rot), None))
    else:
        rot = np.ascontiguousarray(obj.transpose(2, 0, 1, 3))
    probe = np.ascontiguousarray(np.stack([probe_real, probe_imag], -1), dtype='float32')
    ppp = 0
    if shift is not None:
        shift = np.ascontiguousarray(shift, dtype='float32')
        pb = np.zeros((B, M, Pp, Pp, 2), 'float32')
        ck(L.adm_probe_shift_fwd(B, M, Pp, P(probe), P(shift), P(pb), None))
        probe_in, ppp = pb, 1
    else:
        probe_in = probe
    Hn = np.ascontiguousarray(np.stack([H.real, H.imag], -1), dtype='float32')
    Hp = np.zeros_like(Hn)
    Hc = np.zeros_like(Hn)
    ck(L.adm_prepare_transfer_function(Pp, P(Hn), 0, 1.0 / (Pp * Pp), P(Hp), None))
    ck(L.adm_prepare_transfer_function(Pp, P(Hn), 1, 1.0 / (Pp * Pp), P(Hc), None))
    hsep = 0
    if separable:
        hsep = 1
        hy = H[:, 0].copy()
        hx = H[0, :] / H[0, 0]
        assert np.abs(np.outer(hy, hx) - H).max() < 1e-9
        f = lambda v: np.ascontiguousarray(np.stack([v.real, v.imag], -1), dtype='float32')
        Hp = np.zeros((2, Pp, 2), 'float32')
        Hc = np.zeros((2, Pp, 2), 'float32')
        ck(L.adm_prepare_transfer_function_separable(Pp, P(f(hy)), P(f(hx)), 0, 1.0 / (Pp * Pp), P(Hp), None))
        ck(L.adm_prepare_transfer_function_separable(Pp, P(f(hy)), P(f(hx)), 1, 1.0 / (Pp * Pp), P(Hc), None))
    hps = 0
    if h_gaps is not None:
        # sparse multislice: one separable transfer function per gap, [Nz-1][2][P][2]
        hsep, hps = 1, 1
        f = lambda v: np.ascontiguousarray(np.stack([v.real, v.imag], -1), dtype='float32')
        Hp = np.zeros((len(h_gaps), 2, Pp, 2), 'float32')
        Hc = np.zeros((len(h_gaps), 2, Pp, 2), 'float32')
        for i, hg in enumerate(h_gaps):
            hy, hx = hg[:, 0].copy(), hg[0, :] / hg[0, 0]
            assert np.abs(np.outer(hy, hx) - hg).max() < 1e-9
            ck(L.adm_prepare_transfer_function_separable(Pp, P(f(hy)), P(f(hx)), 0, 1.0 / (Pp * Pp), P(Hp[i]), None))
            ck(L.adm_prepare_transfer_function_separable(Pp, P(f(hy)), P(f(hx)), 1, 1.0 / (Pp * Pp), P(Hc[i]), None))
    Hfp = Hfc = None
    hfsep = 0
    fp = {'none': 0, 'inf': 1, 'near': 2}[free_prop]
    if fp == 2:
        hfsep = int(np.abs(np.outer(Hfree[:, 0], Hfree[0, :] / Hfree[0, 0]) - Hfree).max() < 1e-9)
        Hf = np.ascontiguousarray(np.stack([Hfree.real, Hfree.imag], -1), dtype='float32')
        Hfp, Hfc = np.zeros_like(Hf), np.zeros_like(Hf)
        ck(L.adm_prepare_transfer_function(Pp, P(Hf), 0, 1.0 / (Pp * Pp), P(Hfp), None))
        ck(L.adm_prepare_transfer_function(Pp, P(Hf), 1, 1.0 / (Pp * Pp), P(Hfc), None))
    cfg = MultisliceCfg(P=Pp, n_slices=Z, B=B, M=M, Y=Y, X=X, k1=float(np.float32(k1)), sign_convention=sign_convention,
                        unknown_type=0 if unknown_type == 'delta_beta' else 1, free_prop=fp,
                        normalize_fft=int(normalize_fft), probe_per_pattern=ppp, h_separable=hsep, h_per_slice=hps,
                        hfree_separable=hfsep)
    posi = np.ascontiguousarray(pos, dtype='int32')
    ckpt = np.full((B, M, Z, Pp, Pp + 2, 2), np.nan, 'float32')
    det = np.zeros((B, M, Pp, Pp, 2), 'float32')
    Xp = (X + 1) & ~1
    utype = 0 if unknown_type == 'delta_beta' else 1
    tvol = np.full((Z, Y, Xp, 2), np.nan, 'float32')
    if fused_chain:
        ck(L.adm_object_to_transmission(Y, X, Zfull, P(obj), P(grid), utype, float(np.float32(k1)), sign_convention, binning,
                                        P(tvol), None))
    else:
        ck(L.adm_transmission_fwd(utype, float(np.float32(k1)), sign_convention, Z * Y, X, P(rot), P(tvol), None))
    ws = np.zeros(int(L.adm_multislice_workspace_bytes(Pp)), 'uint8')
    ck(L.adm_multislice_fwd(ctypes.byref(cfg), P(tvol), P(probe_in), P(posi), P(Hp), P(Hfp), P(ckpt), P(det), P(ws), None))
    pred = np.zeros((B, Pp, Pp), 'float32')
    gdet = np.zeros_like(det)
    lpp = np.zeros((B,), 'float32')
    meas = np.ascontiguousarray(prj_abs, dtype='float32')
    mask = None
    n_valid = Pp * Pp
    if beamstop is not None:
        mask = np.ascontiguousarray((np.asarray(beamstop) >= 1e-5).astype('float32'))
        n_valid = int(mask.sum())
    inv_n = 1.0 / (B * n_valid)
    ck(L.adm_detector_loss(B, M, Pp, 0 if loss == 'lsq' else 1, 0 if raw_data_type == 'magnitude' else 1,
                           poisson_multiplier, inv_n, P(det), P(meas), P(mask), P(pred), P(gdet), P(lpp), None))
    loss_val = float(lpp.astype('float64').sum() * inv_n)
    g_tvol = np.zeros((Z, Y + Pp, (X + Pp + 1) & ~1, 2), 'float32')
    g_probe_in = np.zeros_like(probe_in)
    ck(L.adm_multislice_bwd(ctypes.byref(cfg), P(tvol), P(posi), P(Hc), P(Hfc), P(ckpt), P(gdet), P(g_tvol),
                            P(g_probe_in), P(ws), None))
    assert not ws[:4096].any(), 'hand-off flags must be left zero'
    g_rot = None
    if not fused_chain:
        g_rot = np.full(rot.shape, np.nan, 'float32')
        ck(L.adm_transmission_adj(utype, float(np.float32(k1)), sign_convention, Z, Y, X, Pp, P(g_tvol), P(tvol), P(g_rot), None))
    out = {'loss': loss_val, 'pred': pred, 'det': det}
    wl1 = None if weight_l1 is None else np.ascontiguousarray(weight_l1, dtype='float32')
    if shift is not None:
        g_probe = np.zeros_like(probe)
        g_shift = np.zeros((B, 2), 'float32')
        ck(L.adm_probe_shift_bwd(B, M, Pp, P(probe), P(shift), P(g_probe_in), P(g_probe), P(g_shift), None))
        out['g_shift'] = g_shift
    else:
        g_probe = g_probe_in
    out['g_probe_real'] = g_probe[..., 0]
    out['g_probe_imag'] = g_probe[..., 1]
    if fused_chain:
        # the regulariser gradient INITIALISES the buffer (accumulate = 0), the transpose chain accumulates into it, per slab
        g_obj = np.full(obj.shape, np.nan, 'float32')
        one = np.ones((1,), 'float32')
        edges = [int(round(i * Y / grad_slabs)) for i in range(grad_slabs + 1)]
        for y0, y1 in zip(edges[:-1], edges[1:]):
            if y1 <= y0:
                continue
            if reg:
                ck(L.adm_regularizer_l1tv_range(Y, X, Zfull, P(obj), P(wl1), float(reg.get('alpha_d') or 0.),
                                                float(reg.get('alpha_b') or 0.), float(reg.get('gamma') or 0.), None, P(g_obj),
                                                P(one), 0, y0, y1, None))
            else:
                g_obj[y0:y1] = 0
            ck(L.adm_transmission_to_object_adj(Y, X, Zfull, Pp, P(g_tvol), P(tvol), P(grid), utype, float(np.float32(k1)),
                                                sign_convention, binning, P(g_obj), y0, y1, None))
        if reg:
            val = np.zeros((1,), 'float32')
            ck(L.adm_regularizer_l1tv_scaled(Y, X, Zfull, P(obj), P(wl1), float(reg.get('alpha_d') or 0.),
                                             float(reg.get('alpha_b') or 0.), float(reg.get('gamma') or 0.), P(val), None, None,
                                             0, None))
            out['loss'] += float(val[0])
    else:
        if grid is not None:
            g_obj = np.zeros_like(obj)
            ck(L.adm_rotate_adj(Y, X, Z, P(g_rot), P(grid), P(g_obj), None))
        else:
            g_obj = np.ascontiguousarray(g_rot.transpose(1, 2, 0, 3))
        if reg:
            val = np.zeros((1,), 'float32')
            ck(L.adm_regularizer_l1tv(Y, X, Z, P(obj), P(wl1), float(reg.get('alpha_d') or 0.), float(reg.get('alpha_b') or 0.),
                                      float(reg.get('gamma') or 0.), P(val), P(g_obj), None))
            out['loss'] += float(val[0])
    out['g_obj'] = g_obj
    return out


def multidist_loss_and_grads_emu(obj, probe_real, probe_imag, prj_abs, h_list, k1, *, unknown_type='delta_beta',
                                 loss='lsq', raw_data_type='magnitude', shift=None):
    """One MultiDistModel loss + gradient evaluation on the host-emulated large-frame kernels, mirroring
    adorym_b200.ops.MultiDistFn / FrameShiftFn / DetectorLossLargeFn."""
    from adorym_b200._lib import MultiDistCfg
    L = emu()
    N = obj.shape[0]
    M = probe_real.shape[0]
    D = len(h_list)
    order = np.zeros(N, 'int32')
    ck(L.adm_lf_spectrum_order(N, P(order)))
    hp = np.stack([np.stack([h[np.ix_(order, order)].real, h[np.ix_(order, order)].imag], -1) for h in h_list])
    Hp = np.ascontiguousarray(hp.astype('float32') * np.float32(1.0 / (N * N)))
    Hc = np.ascontiguousarray(Hp * np.array([1, -1], 'float32'))
    cfg = MultiDistCfg(N=N, M=M, D=D, k1=float(np.float32(k1)), sign_convention=1,
                       unknown_type=0 if unknown_type == 'delta_beta' else 1)
    o2 = np.ascontiguousarray(obj.reshape(N, N, 2), dtype='float32')
    probe = np.ascontiguousarray(np.stack([probe_real, probe_imag], -1), dtype='float32')
    work1 = np.full((M, N, N, 2), np.nan, 'float32')
    det = np.full((D, M, N, N, 2), np.nan, 'float32')
    ck(L.adm_multidist_fwd(ctypes.byref(cfg), P(o2), P(probe), P(Hp), P(work1), P(det), None))
    meas = np.ascontiguousarray(prj_abs, dtype='float32')
    out = {}
    if shift is not None:
        shift = np.ascontiguousarray(shift, dtype='float32')
        work = np.full((D, N, N, 2), np.nan, 'float32')
        spec = np.full((D, N, N, 2), np.nan, 'float32')
        shifted = np.full((D, N, N), np.nan, 'float32')
        ck(L.adm_frame_shift_fwd(N, D, P(meas), P(shift), P(order), P(work), P(spec), P(shifted), None))
        out['shifted'] = shifted
        meas_in = shifted
    else:
        meas_in = meas
    pred = np.zeros((D, N, N), 'float32')
    g_det = np.zeros_like(det)
    g_meas = np.zeros((D, N, N), 'float32')
    partial = np.zeros((D, 64), 'float32')
    lpp = np.zeros((D,), 'float32')
    inv_n = 1.0 / (D * N * N)
    ck(L.adm_detector_loss_large(D, M, N * N, 0 if loss == 'lsq' else 1, 0 if raw_data_type == 'magnitude' else 1, 1.0,
                                 inv_n, P(det), P(meas_in), P(pred), P(g_det), P(g_meas), P(partial), P(lpp), None))
    out['loss'] = float(lpp.astype('float64').sum() * inv_n)
    out['pred'] = pred
    work2 = np.full((D, M, N, N, 2), np.nan, 'float32')
    g_obj = np.full((N, N, 2), np.nan, 'float32')
    g_probe = np.full((M, N, N, 2), np.nan, 'float32')
    ck(L.adm_multidist_bwd(ctypes.byref(cfg), P(o2), P(probe), P(Hc), P(g_det), P(work1), P(work2), P(g_obj), P(g_probe), None))
    out['g_obj'] = g_obj.reshape(obj.shape)
    out['g_probe_real'] = g_probe[..., 0]
    out['g_probe_imag'] = g_probe[..., 1]
    if shift is not None:
        g_shift = np.zeros((D, 2), 'float32')
        ck(L.adm_frame_shift_bwd(N, D, P(g_meas), P(spec), P(order), P(work), P(g_shift), None))
        out['g_shift'] = g_shift
    return out
