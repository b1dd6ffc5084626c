"""torch.autograd.Function wrappers over the C ABI (include/adorym_b200.h).

These replace, on the hot path, the reference's `wrappers.py` fft2/ifft2/exp_complex/grid_sample calls and the
autograd-traced slice loop of `propagate.py:208-258` by hand-written sm_100a kernels with analytic adjoints.
All tensors are contiguous fp32 CUDA tensors; complex fields are [..., 2] (re, im).  No CPU path exists.
"""
import ctypes
import math
import numpy as np
import torch

from . import _lib
from ._lib import MultisliceCfg, call, ptr, stream_ptr  # noqa: F401  (re-exported: dist.PeerUpdate uses ops.call / ops.ptr)

PI = 3.14159265359          # adorym/constants.py
UNKNOWN_TYPES = {'delta_beta': 0, 'real_imag': 1}
LOSS_TYPES = {'lsq': 0, 'poisson': 1}
RAW_TYPES = {'magnitude': 0, 'intensity': 1}
SUPPORTED_P = (64, 72)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------------
# transfer functions
# ------------------------------------------------------------------------------------------------------
def prepare_transfer_function(h, device, conj=False):
    """h: complex ndarray [P, P] in natural fftfreq order (adorym/propagate.py:62-81 get_kernel, float64).
    Returns the permuted / transposed / (1/P^2)-scaled fp32 device copy the FFT passes read."""
    h = np.asarray(h)
    P = h.shape[0]
    if h.shape != (P, P) or P not in SUPPORTED_P:
        raise NotImplementedError('adorym_b200: transfer function of shape %s (supported P: %s)' % (h.shape, SUPPORTED_P))
    hn = torch.from_numpy(np.stack([h.real, h.imag], -1).astype('float32')).to(device)   # f64 -> f32 as propagate.py:202-204
    out = torch.empty_like(hn)
    call('adm_prepare_transfer_function', P, ptr(hn), int(bool(conj)), 1.0 / (P * P), ptr(out), stream_ptr())
    return out


def separate_transfer_function(h, rtol=1e-6):
    """If h[ky, kx] == hy[ky] * hx[kx] (true for the paraxial Fresnel kernel, propagate.py:72) return (hy, hx) as
    complex128 vectors, else None."""
    h = np.asarray(h, dtype='complex128')
    if h.ndim != 2 or abs(h[0, 0]) < 1e-30:
        return None
    hy = h[:, 0].copy()
    hx = h[0, :] / h[0, 0]
    if np.abs(np.outer(hy, hx) - h).max() > rtol * np.abs(h).max():
        return None
    return hy, hx


def prepare_transfer_function_separable(hy, hx, device, conj=False):
    P = hy.shape[0]
    if P not in SUPPORTED_P:
        raise NotImplementedError('adorym_b200: transfer function of size %d (supported P: %s)' % (P, SUPPORTED_P))
    f = lambda v: torch.from_numpy(np.stack([v.real, v.imag], -1).astype('float32')).to(device)
    hyt, hxt = f(hy), f(hx)
    out = torch.empty((2, P, 2), dtype=torch.float32, device=device)
    call('adm_prepare_transfer_function_separable', P, ptr(hyt), ptr(hxt), int(bool(conj)), 1.0 / (P * P), ptr(out),
         stream_ptr())
    return out


class MultislicePlan(object):
    """Static description of one multislice configuration + prepared transfer functions (cached per model)."""

    def __init__(self, P, n_slices, Y, X, k1, h, device, sign_convention=1, unknown_type='delta_beta', free_prop='none',
                 normalize_fft=False, h_free=None, allow_separable=True, h_gaps=None, binning=1, obj_slices=None,
                 lmbda_nm=None, voxel_nm=None):
        """h: the in-object transfer function [P,P] (all gaps equal), or h_gaps: a list of n_slices-1 separable transfer
        functions, one per gap between consecutive slices (sparse multislice, propagate.py:479-534)."""
        if P not in SUPPORTED_P:
            raise NotImplementedError('adorym_b200: probe size %d is not supported yet (supported: %s)' % (P, SUPPORTED_P))
        self.P, self.n_slices, self.Y, self.X = int(P), int(n_slices), int(Y), int(X)
        self.lmbda_nm, self.voxel_nm = lmbda_nm, voxel_nm          # needed only for refinable slice positions
        # binning (propagate.py:207-254): n_slices counts PROPAGATION steps = ceil(obj_slices / binning); each step's
        # transmission is that of (delta, beta) summed over `binning` object slices (layout 'yxz' only)
        self.binning = int(binning)
        self.obj_slices = int(obj_slices) if obj_slices is not None else int(n_slices)
        if self.binning < 1 or -(-self.obj_slices // self.binning) != self.n_slices:
            raise ValueError('n_slices must equal ceil(obj_slices / binning)')
        if self.binning > 1 and unknown_type != 'delta_beta':
            raise NotImplementedError('adorym_b200: binning > 1 is accelerated for delta_beta objects only')
        self.k1 = float(k1)
        self.sign_convention = int(sign_convention)
        self.unknown_type = UNKNOWN_TYPES[unknown_type]
        self.free_prop = {'none': 0, 'inf': 1, 'near': 2}[free_prop]
        self.normalize_fft = int(bool(normalize_fft))
        self.device = device
        self.Hp = self.Hc = self.Hfp = self.Hfc = None
        self.h_separable = 0
        self.h_per_slice = 0
        self.hfree_separable = 0
        if h_gaps is not None:
            if len(h_gaps) != self.n_slices - 1:
                raise ValueError('h_gaps must hold n_slices - 1 transfer functions')
            seps = [separate_transfer_function(hg) for hg in h_gaps]
            if any(sp is None for sp in seps):
                raise NotImplementedError('adorym_b200: per-gap transfer functions must be separable (paraxial Fresnel kernel)')
            self.h_separable, self.h_per_slice = 1, 1
            if seps:
                self.Hp = torch.stack([prepare_transfer_function_separable(a, b, device, conj=False) for a, b in seps]).contiguous()
                self.Hc = torch.stack([prepare_transfer_function_separable(a, b, device, conj=True) for a, b in seps]).contiguous()
        elif h is not None:
            sep = separate_transfer_function(h) if allow_separable else None
            if sep is not None:
                self.h_separable = 1
                self.Hp = prepare_transfer_function_separable(sep[0], sep[1], device, conj=False)
                self.Hc = prepare_transfer_function_separable(sep[0], sep[1], device, conj=True)
            else:
                self.Hp = prepare_transfer_function(h, device, conj=False)
                self.Hc = prepare_transfer_function(h, device, conj=True)
        elif self.n_slices > 1:
            raise ValueError('a transfer function is required when n_slices > 1')
        if self.free_prop == 2:
            self.set_free_kernel(h_free)

    def set_free_kernel(self, h_free):
        self.Hfp = prepare_transfer_function(h_free, self.device, conj=False)
        self.Hfc = prepare_transfer_function(h_free, self.device, conj=True)
        # an outer product hy(ky) hx(kx) (the paraxial kernel always is): the three-pass kernels take the factors from the
        # first row / column of the prepared table
        self.hfree_separable = int(separate_transfer_function(h_free) is not None)

    def cfg(self, B, M, probe_per_pattern):
        return MultisliceCfg(P=self.P, n_slices=self.n_slices, B=int(B), M=int(M), Y=self.Y, X=self.X, k1=self.k1,
                             sign_convention=self.sign_convention, unknown_type=self.unknown_type,
                             free_prop=self.free_prop, normalize_fft=self.normalize_fft,
                             probe_per_pattern=int(bool(probe_per_pattern)), h_separable=self.h_separable,
                             h_per_slice=self.h_per_slice, hfree_separable=self.hfree_separable)


# ------------------------------------------------------------------------------------------------------
# transmission volume + multislice
# ------------------------------------------------------------------------------------------------------
def tvol_pitch(X):
    """ADM_TVOL_PITCH: row pitch of the transmission volume (TMA needs a multiple of 16 bytes)."""
    return (int(X) + 1) & ~1


def gvol_pitch(X, P):
    """ADM_GVOL_PITCH: row pitch of the margined gradient volume dL/dT."""
    return (int(X) + int(P) + 1) & ~1


_MS_WORKSPACE = {}


def multislice_workspace(device, P):
    """Persistent, zero-initialised scratch of the balanced multislice schedule (include/adorym_b200.h:
    adm_multislice_workspace_bytes), one per (device, stream, P): hand-off slots + flags for wavefields that migrate
    between CTAs.  The kernels leave the flags zero, so it is allocated and cleared once."""
    key = (str(device), int(torch.cuda.current_stream(device).cuda_stream), int(P))
    ws = _MS_WORKSPACE.get(key)
    if ws is None:
        n = int(_lib.get_lib().adm_multislice_workspace_bytes(int(P)))
        ws = torch.zeros((n,), dtype=torch.uint8, device=device)
        _MS_WORKSPACE[key] = ws
    return ws


# Gradient slabs: with GRAD_SLABS > 1 the object gradient of MultisliceFn.backward is produced in that many y slabs (the
# rotation is about y, so the planes are independent) and a CUDA event is recorded after each; the driver starts the
# NCCL all-reduce of a slab as soon as its event has fired (dist.Comm.allreduce_slabs) instead of after the whole volume.
GRAD_SLABS = 1
last_obj_grad_slabs = None       # [(y0, y1, event)] of the most recent backward, or None
# Optional preallocated output of the object gradient [Y,X,Z,2] (e.g. the symmetric, peer-mapped gradient buffer of
# dist.PeerUpdate): MultisliceFn.backward then writes the gradient there instead of into a fresh tensor.
OBJ_GRAD_OUT = None


def slab_ranges(Y, n):
    n = max(1, min(int(n), int(Y)))
    edges = [int(round(i * Y / n)) for i in range(n + 1)]
    return [(edges[i], edges[i + 1]) for i in range(n) if edges[i + 1] > edges[i]]


class MultisliceFn(torch.autograd.Function):
    """det[B,M,P,P,2] (, reg) = multislice(obj, probe[M,P,P,2] | [B,M,P,P,2], pos[B,2] int32).

    layout 'yxz': obj is the reference-layout object [Y,X,Z,2]; `grid` (normalized_grid) rotates it about y, grid=None
                  means it already is in the beam frame.  One fused kernel makes the slice-major transmission volume
                  T' = T - 1 (adm_object_to_transmission: util.py:536-552 + propagate.py:241/246), and one fused kernel
                  takes dL/dT back to the object (adm_transmission_to_object_adj).  `reg` = None or
                  (weight_l1|None, alpha_d, alpha_b, gamma): the L1 / TV regulariser of the object
                  (regularizers.py:30-110); its value is returned as a second output and its gradient INITIALISES the
                  object-gradient buffer that the transpose rotation then accumulates into (no zero fill, no add).
    layout 'zyx': obj is an explicit slice-major volume [Nz,Y,X,2] (propagate.py:131 drop-in); reg must be None.

    forward : transmission volume -> adm_multislice_fwd
    backward: adm_multislice_bwd (dL/dT scattered into the margined gradient volume by TMA reduce-adds) -> chain rule."""

    @staticmethod
    def forward(ctx, obj, grid, probe, pos, plan, reg=None, layout='zyx', slice_pos=None):
        """slice_pos: optional tensor [Nz] of refinable slice positions (cm) of the sparse multislice model; the plan's
        per-gap transfer functions must have been prepared from its current values.  It only enters the backward pass
        (dL/d slice_pos, propagate.py:479-534 with fresnel_propagate_wrapped :556-568)."""
        obj = _f32c(obj)
        probe = _f32c(probe)
        if pos.dtype != torch.int32 or not pos.is_cuda:
            raise RuntimeError('pos must be an int32 CUDA tensor [B, 2]')
        pos = pos.contiguous()
        P, Nz, Y, X = plan.P, plan.n_slices, plan.Y, plan.X
        ctx.want_slice_pos = slice_pos is not None and slice_pos.requires_grad
        if ctx.want_slice_pos and not plan.h_per_slice:
            raise RuntimeError('refinable slice positions need a plan with per-gap transfer functions (h_gaps)')
        Zo = plan.obj_slices
        want = (Y, X, Zo, 2) if layout == 'yxz' else (Nz, Y, X, 2)
        if layout != 'yxz' and plan.binning != 1:
            raise RuntimeError('binning needs the reference layout [Y,X,Z,2]')
        if tuple(obj.shape) != want:
            raise RuntimeError('object must be %s for layout %s, got %s' % (want, layout, tuple(obj.shape)))
        if layout != 'yxz' and (reg is not None or grid is not None):
            raise RuntimeError('rotation / fused regulariser need the reference layout [Y,X,Z,2]')
        dev = obj.device
        B = pos.shape[0]
        per_pattern = probe.dim() == 5
        M = probe.shape[1] if per_pattern else probe.shape[0]
        if tuple(probe.shape[-3:]) != (P, P, 2) or (per_pattern and probe.shape[0] != B):
            raise RuntimeError('probe must be [M,P,P,2] or [B,M,P,P,2]')
        cfg = plan.cfg(B, M, per_pattern)
        tvol = torch.empty((Nz, Y, tvol_pitch(X), 2), dtype=torch.float32, device=dev)
        if layout == 'yxz':
            call('adm_object_to_transmission', Y, X, Zo, ptr(obj), ptr(grid), plan.unknown_type, plan.k1,
                 plan.sign_convention, plan.binning, ptr(tvol), stream_ptr())
        else:
            call('adm_transmission_fwd', plan.unknown_type, plan.k1, plan.sign_convention, Nz * Y, X, ptr(obj), ptr(tvol),
                 stream_ptr())
        reg_val = None
        if reg is not None:
            w, ad, ab, gm = reg
            w = _f32c(w) if w is not None else None
            reg = (w, float(ad or 0.), float(ab or 0.), float(gm or 0.))
            val = torch.zeros((1,), dtype=torch.float32, device=dev)
            call('adm_regularizer_l1tv_scaled', Y, X, Zo, ptr(obj), ptr(w), reg[1], reg[2], reg[3], ptr(val), None, None, 0,
                 stream_ptr())
            reg_val = val[0]
        ckpt = torch.empty((B, M, Nz, P, P + 2, 2), dtype=torch.float32, device=dev)
        det = torch.empty((B, M, P, P, 2), dtype=torch.float32, device=dev)
        call('adm_multislice_fwd', ctypes.byref(cfg), ptr(tvol), ptr(probe), ptr(pos), ptr(plan.Hp), ptr(plan.Hfp),
             ptr(ckpt), ptr(det), ptr(multislice_workspace(dev, P)), stream_ptr())
        ctx.plan, ctx.cfg, ctx.per_pattern, ctx.layout, ctx.reg = plan, cfg, per_pattern, layout, reg
        ctx.probe_shape = tuple(probe.shape)
        ctx.has_grid = grid is not None
        saved = [tvol, pos, ckpt]
        if grid is not None:
            saved.append(grid)
        if reg is not None:
            saved.append(obj)
            if reg[0] is not None:
                saved.append(reg[0])
        ctx.save_for_backward(*saved)
        if reg is not None:
            return det, reg_val
        return det

    @staticmethod
    def backward(ctx, g_det, g_reg=None):
        global last_obj_grad_slabs
        saved = list(ctx.saved_tensors)
        tvol, pos, ckpt = saved[:3]
        rest = saved[3:]
        grid = rest.pop(0) if ctx.has_grid else None
        obj = rest.pop(0) if ctx.reg is not None else None
        w = rest.pop(0) if (ctx.reg is not None and ctx.reg[0] is not None) else None
        plan = ctx.plan
        P, Nz, Y, X = plan.P, plan.n_slices, plan.Y, plan.X
        Zo = plan.obj_slices
        dev = tvol.device
        g_det = _f32c(g_det)
        # margined gradient volume (ADM_GVOL_MARGIN in include/adorym_b200.h); the adjoint kernel reduce-adds into it
        g_tvol = torch.zeros((Nz, Y + P, gvol_pitch(X, P), 2), dtype=torch.float32, device=dev)
        g_probe = None
        if ctx.needs_input_grad[2]:
            g_probe = torch.zeros(ctx.probe_shape, dtype=torch.float32, device=dev)
        g_slice_pos = None
        if ctx.want_slice_pos:
            B, M = ckpt.shape[0], ckpt.shape[1]
            g_psi = torch.empty((B, M, Nz, P, P, 2), dtype=torch.float32, device=dev)
            call('adm_multislice_bwd_gpsi', ctypes.byref(ctx.cfg), ptr(tvol), ptr(pos), ptr(plan.Hc), ptr(plan.Hfc), ptr(ckpt),
                 ptr(g_det), ptr(g_tvol), ptr(g_probe), ptr(g_psi), ptr(multislice_workspace(dev, P)), stream_ptr())
            g_slice_pos = _slice_pos_gradient(plan, ckpt, g_psi, pos)
        else:
            call('adm_multislice_bwd', ctypes.byref(ctx.cfg), ptr(tvol), ptr(pos), ptr(plan.Hc), ptr(plan.Hfc), ptr(ckpt),
                 ptr(g_det), ptr(g_tvol), ptr(g_probe), ptr(multislice_workspace(dev, P)), stream_ptr())
        g_obj = None
        if ctx.needs_input_grad[0]:
            if ctx.layout == 'yxz':
                if OBJ_GRAD_OUT is not None and tuple(OBJ_GRAD_OUT.shape) == (Y, X, Zo, 2) and OBJ_GRAD_OUT.device == dev:
                    g_obj = OBJ_GRAD_OUT
                else:
                    g_obj = torch.empty((Y, X, Zo, 2), dtype=torch.float32, device=dev)
                use_reg = ctx.reg is not None and g_reg is not None
                gs = _f32c(g_reg.reshape(1)) if use_reg else None
                slabs = []
                ranges = slab_ranges(Y, GRAD_SLABS)
                # the regulariser gradient INITIALISES the buffer: for all planes first (it is the last reader of the object
                # in a step -- with the pipelined peer-memory update a finished slab may be overwritten by the peers while
                # later slabs are still being produced), then the transpose chain slab by slab
                if use_reg:
                    call('adm_regularizer_l1tv_range', Y, X, Zo, ptr(obj), ptr(w), ctx.reg[1], ctx.reg[2], ctx.reg[3], None,
                         ptr(g_obj), ptr(gs), 0, 0, Y, stream_ptr())
                else:
                    g_obj.zero_()
                for (y0, y1) in ranges:
                    call('adm_transmission_to_object_adj', Y, X, Zo, P, ptr(g_tvol), ptr(tvol), ptr(grid), plan.unknown_type,
                         plan.k1, plan.sign_convention, plan.binning, ptr(g_obj), y0, y1, stream_ptr())
                    if GRAD_SLABS > 1:
                        ev = torch.cuda.Event()
                        ev.record()
                        slabs.append((y0, y1, ev))
                last_obj_grad_slabs = (g_obj, slabs) if slabs else None
            else:
                g_obj = torch.empty((Nz, Y, X, 2), dtype=torch.float32, device=dev)
                call('adm_transmission_adj', plan.unknown_type, plan.k1, plan.sign_convention, Nz, Y, X, P, ptr(g_tvol),
                     ptr(tvol), ptr(g_obj), stream_ptr())
        return g_obj, None, g_probe, None, None, None, None, g_slice_pos


def _slice_pos_gradient(plan, ckpt, g_psi, pos):
    """dL/d slice_pos [Nz] (per cm) of the sparse multislice model from the wavefields psi_z ENTERING every slice (the
    checkpoints) and the adjoint wavefields g_z = dL/d psi_z: the gap d_i between slices i and i+1 enters through
    psi_{i+1} = IFFT2(H(d_i) FFT2(.)), H = exp(i phi' d_i), so d psi_{i+1} / d d_i = IFFT2(i phi' FFT2(psi_{i+1})) and, by
    Parseval, dL/d d_i = -(1/P^2) sum_k phi'_k Im(conj(G_k) Psi_k) over all wavefields; slice_pos[j] appears in d_{j-1} with +
    and in d_j with - .  A few slices of a few wavefields: library FFTs, not a hot path."""
    B, M, Nz, P = ckpt.shape[0], ckpt.shape[1], ckpt.shape[2], plan.P
    dev = ckpt.device
    # checkpoint rows carry the tile pitch P + 2 at column offset (px & 1)
    off = (pos[:, 1] & 1).long()
    cols = off.view(B, 1) + torch.arange(P, device=dev).view(1, P)                      # [B, P]
    psi = torch.gather(ckpt, 4, cols.view(B, 1, 1, 1, P, 1).expand(B, M, Nz, P, P, 2))
    Psi = torch.fft.fft2(torch.view_as_complex(psi.contiguous())[:, :, 1:])             # psi_{i+1}, i = 0 .. Nz-2
    G = torch.fft.fft2(torch.view_as_complex(g_psi)[:, :, 1:])
    f = torch.fft.fftfreq(P, device=dev, dtype=torch.float32) / float(plan.voxel_nm)
    dphi = (-float(plan.sign_convention) * PI * float(plan.lmbda_nm)) * (f[:, None] ** 2 + f[None, :] ** 2)   # per nm
    im = (torch.conj(G) * Psi).imag.double()
    g_gap = -(im * dphi.double()).sum(dim=(0, 1, 3, 4)) / float(P * P)                  # [Nz - 1], per nm
    g_pos = torch.zeros((Nz,), dtype=torch.float64, device=dev)
    g_pos[1:] += g_gap
    g_pos[:-1] -= g_gap
    return (g_pos * 1e7).to(torch.float32)


# ------------------------------------------------------------------------------------------------------
# rotation
# ------------------------------------------------------------------------------------------------------
def normalized_grid(coord_old, X, Z, device):
    """LUT (fp16 [X*Z, 2], columns x', z') -> normalised grid_sample coordinates (gz, gx) fp32 on `device`, with the
    arithmetic of adorym/wrappers.py:1124-1141 on the CPU path (fp64 upcast at util.py:539-540, cast at :1141)."""
    g = np.asarray(coord_old).astype('float64')[:, ::-1]
    shp32 = np.array([X, Z], dtype='float32')                     # sic: wrappers.py:1135-1137 divides (z', x') by (X, Z)
    inv = (np.float32(1.) / shp32).astype('float64')              # `1. / arr_shape` is an fp32 tensor op there
    g = -1 + 2. * g / shp32.astype('float64') + inv
    return torch.from_numpy(np.ascontiguousarray(g.astype('float32'))).to(device)


class RotateFn(torch.autograd.Function):
    """rot[Z,Y,X,2] (slice-major) = bilinear/border rotation of obj[Y,X,Z,2] about y given the normalised grid."""

    @staticmethod
    def forward(ctx, obj, grid):
        obj = _f32c(obj)
        Y, X, Z, _ = obj.shape
        rot = torch.empty((Z, Y, X, 2), dtype=torch.float32, device=obj.device)
        call('adm_rotate_fwd', Y, X, Z, ptr(obj), ptr(grid), ptr(rot), stream_ptr())
        ctx.save_for_backward(grid)
        ctx.shape = (Y, X, Z)
        return rot

    @staticmethod
    def backward(ctx, g_rot):
        (grid,) = ctx.saved_tensors
        Y, X, Z = ctx.shape
        g_obj = torch.zeros((Y, X, Z, 2), dtype=torch.float32, device=g_rot.device)
        call('adm_rotate_adj', Y, X, Z, ptr(_f32c(g_rot)), ptr(grid), ptr(g_obj), stream_ptr())
        return g_obj, None


# ------------------------------------------------------------------------------------------------------
# detector magnitude / loss
# ------------------------------------------------------------------------------------------------------
def _detector_call(det, meas, mask, loss_type, raw_type, pm, inv_n, want_grad):
    B, M, P = det.shape[0], det.shape[1], det.shape[2]
    pred = torch.empty((B, P, P), dtype=torch.float32, device=det.device)
    lpp = torch.empty((B,), dtype=torch.float32, device=det.device)
    g_det = torch.empty_like(det) if want_grad else None
    call('adm_detector_loss', B, M, P, loss_type, raw_type, float(pm), float(inv_n), ptr(det), ptr(meas), ptr(mask),
         ptr(pred), ptr(g_det), ptr(lpp), stream_ptr())
    return pred, lpp, g_det


class DetectorLossFn(torch.autograd.Function):
    """Fused |det| + mismatch loss + residual.  Returns (loss, pred[B,P,P], loss_per_pattern[B])."""

    @staticmethod
    def forward(ctx, det, meas, mask, loss_function_type, raw_data_type, poisson_multiplier):
        det = _f32c(det)
        meas = _f32c(meas)
        B, M, P = det.shape[0], det.shape[1], det.shape[2]
        n_valid = P * P if mask is None else int((mask >= 1e-5).sum().item())
        inv_n = 1.0 / (B * n_valid)
        pred, lpp, g_det = _detector_call(det, meas, mask, LOSS_TYPES[loss_function_type], RAW_TYPES[raw_data_type],
                                          poisson_multiplier, inv_n, True)
        ctx.save_for_backward(g_det)
        loss = lpp.sum() * inv_n
        ctx.mark_non_differentiable(pred, lpp)
        return loss, pred, lpp

    @staticmethod
    def backward(ctx, g_loss, g_pred, g_lpp):
        (g_det,) = ctx.saved_tensors
        return g_det * g_loss, None, None, None, None, None


class DetectorMagnitudeFn(torch.autograd.Function):
    """pred[B,P,P] = sqrt(sum_m |det_m|^2) (forward_model.py:353 / :371-375), zero subgradient at 0."""

    @staticmethod
    def forward(ctx, det):
        det = _f32c(det)
        B, M, P = det.shape[0], det.shape[1], det.shape[2]
        dummy = torch.zeros((B, P, P), dtype=torch.float32, device=det.device)
        pred, _, _ = _detector_call(det, dummy, None, 0, 0, 1.0, 1.0, False)
        ctx.save_for_backward(det, pred)
        return pred

    @staticmethod
    def backward(ctx, g_pred):
        det, pred = ctx.saved_tensors
        c = torch.where(pred > 0, g_pred / pred, torch.zeros_like(pred))
        return det * c[:, None, :, :, None]


# ------------------------------------------------------------------------------------------------------
# probe shift
# ------------------------------------------------------------------------------------------------------
class ProbeShiftFn(torch.autograd.Function):
    """out[B,M,P,P,2] = per-pattern Fourier shift of probe[M,P,P,2] by shift[B,2] (y, x) (util.py:380-397)."""

    @staticmethod
    def forward(ctx, probe, shift):
        probe = _f32c(probe)
        shift = _f32c(shift)
        M, P = probe.shape[0], probe.shape[1]
        B = shift.shape[0]
        out = torch.empty((B, M, P, P, 2), dtype=torch.float32, device=probe.device)
        call('adm_probe_shift_fwd', B, M, P, ptr(probe), ptr(shift), ptr(out), stream_ptr())
        ctx.save_for_backward(probe, shift)
        return out

    @staticmethod
    def backward(ctx, g_out):
        probe, shift = ctx.saved_tensors
        M, P = probe.shape[0], probe.shape[1]
        B = shift.shape[0]
        g_probe = torch.zeros_like(probe)
        g_shift = torch.zeros_like(shift)
        call('adm_probe_shift_bwd', B, M, P, ptr(probe), ptr(shift), ptr(_f32c(g_out)), ptr(g_probe), ptr(g_shift),
             stream_ptr())
        return g_probe, g_shift


# ------------------------------------------------------------------------------------------------------
# regulariser / Adam
# ------------------------------------------------------------------------------------------------------
class RegL1TVFn(torch.autograd.Function):
    """alpha_d mean(w0|d|) + alpha_b mean(w1|b|) + gamma (TV3d(d) + TV3d(b)) on obj[Y,X,Z,2] (delta_beta).

    forward : value only (one read of the object).
    backward: the gradient, already multiplied by the upstream gradient (read on the device, no host sync), written
              in one pass -- no zero fill, no separate scaling kernel."""

    @staticmethod
    def forward(ctx, obj, weight_l1, alpha_d, alpha_b, gamma):
        obj = _f32c(obj)
        Y, X, Z, _ = obj.shape
        val = torch.zeros((1,), dtype=torch.float32, device=obj.device)
        w = _f32c(weight_l1) if weight_l1 is not None else None
        ctx.coef = (float(alpha_d or 0.), float(alpha_b or 0.), float(gamma or 0.))
        call('adm_regularizer_l1tv_scaled', Y, X, Z, ptr(obj), ptr(w), ctx.coef[0], ctx.coef[1], ctx.coef[2], ptr(val), None,
             None, 0, stream_ptr())
        ctx.has_w = w is not None
        if w is not None:
            ctx.save_for_backward(obj, w)
        else:
            ctx.save_for_backward(obj)
        return val[0]

    @staticmethod
    def backward(ctx, g_val):
        obj = ctx.saved_tensors[0]
        w = ctx.saved_tensors[1] if ctx.has_w else None
        Y, X, Z, _ = obj.shape
        g = torch.empty_like(obj)
        gs = _f32c(g_val.reshape(1))
        call('adm_regularizer_l1tv_scaled', Y, X, Z, ptr(obj), ptr(w), ctx.coef[0], ctx.coef[1], ctx.coef[2], None, ptr(g),
             ptr(gs), 0, stream_ptr())
        return g, None, None, None, None


def adam_step_(x, m, v, g, step, step_size=0.001, b1=0.9, b2=0.999, eps=1e-7, clip_nonneg=False,
               zero_channel=None, mask=None):
    """In-place fused Adam on fp32 CUDA tensors (optimizers.py:309-318) + optional epilogue."""
    for name, t in (('x', x), ('m', m), ('v', v), ('g', g)):
        if t.dtype != torch.float32:
            raise RuntimeError('adam_step_: fp32 tensors required (%s is %s)' % (name, t.dtype))
        if t.device != x.device or t.numel() != x.numel():
            raise RuntimeError('adam_step_: %s must live on %s and hold %d elements (got %s, %d)'
                               % (name, x.device, x.numel(), t.device, t.numel()))
    for name, t in (('x', x), ('m', m), ('v', v)):
        if not t.is_contiguous():
            raise RuntimeError('adam_step_: %s must be contiguous (it is updated in place)' % name)
    if mask is not None:
        # the reference multiplies obj [..., 2] by mask [..., 1] (array_ops.py:239-244): a broadcast error there, a
        # silent out-of-bounds read here unless checked
        if mask.dtype != torch.float32 or mask.device != x.device or not mask.is_contiguous() \
                or mask.numel() * 2 != x.numel():
            raise RuntimeError('adam_step_: mask must be a contiguous fp32 tensor on %s with %d elements (one per voxel); got '
                               '%s %s %s' % (x.device, x.numel() // 2, mask.dtype, mask.device, tuple(mask.shape)))
    flags = (1 if clip_nonneg else 0) | (2 if zero_channel == 0 else 0) | (4 if zero_channel == 1 else 0)
    call('adm_adam_step', ptr(x), ptr(m), ptr(v), ptr(_f32c(g)), x.numel(), float(step_size), float(b1), float(b2),
         float(eps), int(step), flags, ptr(mask), stream_ptr())
    return x


def sgd_step_(x, v, g, step_size=0.001, gamma=0.9, clip_nonneg=False, zero_channel=None, mask=None):
    """In-place gradient descent (v is None) or momentum update on fp32 CUDA tensors (optimizers.py:366-464)."""
    for name, t in (('x', x), ('g', g)) + ((('v', v),) if v is not None else ()):
        if t.dtype != torch.float32 or t.device != x.device or t.numel() != x.numel():
            raise RuntimeError('sgd_step_: %s must be fp32 on %s with %d elements' % (name, x.device, x.numel()))
    if not x.is_contiguous() or (v is not None and not v.is_contiguous()):
        raise RuntimeError('sgd_step_: x and v must be contiguous (they are updated in place)')
    if mask is not None and (mask.dtype != torch.float32 or mask.device != x.device or not mask.is_contiguous()
                             or mask.numel() * 2 != x.numel()):
        raise RuntimeError('sgd_step_: mask must be a contiguous fp32 tensor with one element per voxel')
    flags = (1 if clip_nonneg else 0) | (2 if zero_channel == 0 else 0) | (4 if zero_channel == 1 else 0)
    call('adm_sgd_step', ptr(x), ptr(v), ptr(_f32c(g)), x.numel(), float(step_size), float(gamma), flags, ptr(mask),
         stream_ptr())
    return x


def k1_of(energy_ev, psize_cm, scale_ri_by_k=True, ds_level=1):
    """k1 = 2 pi dz / lambda (propagate.py:148,215)."""
    if not scale_ri_by_k:
        return 1.0
    lmbda_nm = 1240. / energy_ev
    return 2. * PI * (psize_cm * ds_level * 1.e7) / lmbda_nm


# ------------------------------------------------------------------------------------------------------
# large frames: multi-distance near-field propagation of one thin object (MultiDistModel) + measured-frame shift
# ------------------------------------------------------------------------------------------------------
_LF_ORDER = {}


def lf_supported(N):
    return bool(_lib.get_lib().adm_lf_supported(int(N)))


def lf_spectrum_order(N, device=None):
    """freq index at each position of the permuted spectrum (host ndarray, or int32 device tensor if device given)."""
    key = (int(N), str(device))
    if key not in _LF_ORDER:
        buf = (ctypes.c_int * int(N))()
        _lib.check(_lib.get_lib().adm_lf_spectrum_order(int(N), ctypes.cast(buf, ctypes.c_void_p)), 'adm_lf_spectrum_order')
        order = np.frombuffer(buf, dtype=np.int32).copy()
        _LF_ORDER[key] = order if device is None else torch.from_numpy(order).to(device)
    return _LF_ORDER[key]


class MultiDistPlan(object):
    """Prepared transfer functions of all distances for N x N frames: Hprep[d][py][px] = H_d[f(py)][f(px)] / N^2
    (adm_lf_spectrum_order), plus its conjugate for the adjoint."""

    def __init__(self, N, h_list, k1, device, sign_convention=1, unknown_type='delta_beta'):
        if not lf_supported(N):
            raise NotImplementedError('adorym_b200: frame size %d is not supported by the large-frame path' % N)
        order = lf_spectrum_order(N)
        hp = []
        for h in h_list:
            h = np.asarray(h)
            if h.shape != (N, N):
                raise ValueError('transfer function shape %s != (%d, %d)' % (h.shape, N, N))
            hh = h[np.ix_(order, order)]
            hp.append(np.stack([hh.real, hh.imag], -1).astype('float32'))      # f64 -> f32 as propagate.py:202-204
        hp = torch.from_numpy(np.stack(hp)).to(device) * (1.0 / (N * N))
        self.Hp = hp.contiguous()
        self.Hc = (hp * torch.tensor([1.0, -1.0], device=device)).contiguous()
        self.N, self.D = int(N), len(h_list)
        self.k1, self.sign_convention, self.unknown_type = float(k1), int(sign_convention), UNKNOWN_TYPES[unknown_type]
        self.device = device

    def cfg(self, M):
        return _lib.MultiDistCfg(N=self.N, M=int(M), D=self.D, k1=self.k1, sign_convention=self.sign_convention,
                                 unknown_type=self.unknown_type)


class MultiDistFn(torch.autograd.Function):
    """det[D,M,N,N,2] = IFFT2(H_d . FFT2(probe_m . T(obj))): adm_multidist_fwd / adm_multidist_bwd."""

    @staticmethod
    def forward(ctx, obj2d, probe, plan):
        obj2d = _f32c(obj2d)
        probe = _f32c(probe)
        N, D = plan.N, plan.D
        M = probe.shape[0]
        if tuple(obj2d.shape) != (N, N, 2) or tuple(probe.shape) != (M, N, N, 2):
            raise RuntimeError('MultiDistFn: obj [N,N,2] / probe [M,N,N,2] expected, got %s / %s'
                               % (tuple(obj2d.shape), tuple(probe.shape)))
        dev = obj2d.device
        cfg = plan.cfg(M)
        work1 = torch.empty((M, N, N, 2), dtype=torch.float32, device=dev)
        det = torch.empty((D, M, N, N, 2), dtype=torch.float32, device=dev)
        call('adm_multidist_fwd', ctypes.byref(cfg), ptr(obj2d), ptr(probe), ptr(plan.Hp), ptr(work1), ptr(det), stream_ptr())
        ctx.plan, ctx.cfg = plan, cfg
        ctx.save_for_backward(obj2d, probe)
        return det

    @staticmethod
    def backward(ctx, g_det):
        obj2d, probe = ctx.saved_tensors
        plan = ctx.plan
        N, D = plan.N, plan.D
        M = probe.shape[0]
        dev = obj2d.device
        g_det = _f32c(g_det)
        work1 = torch.empty((M, N, N, 2), dtype=torch.float32, device=dev)
        work2 = torch.empty((D, M, N, N, 2), dtype=torch.float32, device=dev)
        g_obj = torch.empty_like(obj2d) if ctx.needs_input_grad[0] else None
        g_probe = torch.empty_like(probe) if ctx.needs_input_grad[1] else None
        call('adm_multidist_bwd', ctypes.byref(ctx.cfg), ptr(obj2d), ptr(probe), ptr(plan.Hc), ptr(g_det), ptr(work1),
             ptr(work2), ptr(g_obj), ptr(g_probe), stream_ptr())
        return g_obj, g_probe, None


class MultiDistRefineFn(torch.autograd.Function):
    """MultiDistFn with REFINABLE propagation distances (optimize_free_prop: propagate.py:271-276 ->
    fresnel_propagate_wrapped :556-568 -> get_kernel_wrapped :84-103, the paraxial kernel as a differentiable function of
    the distance): det[d] = IFFT2(H(dist_d) . FFT2(probe . T(obj))), H(dist) = exp(-i s pi lambda dist (u^2 + v^2)).

    The transfer functions are rebuilt on the device from the current distances (fp32, like the reference's in-graph H);
    dL/d dist_d = Re sum conj(g_det) . IFFT2(i phi' H_d F) = sum g_det . IFFT2(i phi' FFT2(det_d)), phi' = d phase / d dist:
    one more fused large-frame pass per distance on the detector field itself (vacuum object, transfer function i phi')."""

    @staticmethod
    def forward(ctx, obj2d, probe, dists_cm, geom):
        """geom: dict(N, k1, sign_convention, unknown_type, lmbda_nm, voxel_nm)"""
        obj2d, probe = _f32c(obj2d), _f32c(probe)
        N = int(geom['N'])
        dev = obj2d.device
        M, D = probe.shape[0], int(dists_cm.numel())
        order = lf_spectrum_order(N, dev).long()
        f = torch.fft.fftfreq(N, device=dev, dtype=torch.float32)[order] / float(geom['voxel_nm'])
        u2v2 = (f * f)[:, None] + (f * f)[None, :]
        dphi = (-float(geom['sign_convention']) * PI * float(geom['lmbda_nm'])) * u2v2          # phase per nm of distance
        ph = (dists_cm.detach().to(dev, torch.float32) * 1e7)[:, None, None] * dphi[None]
        Hp = (torch.stack([torch.cos(ph), torch.sin(ph)], -1) * (1.0 / (N * N))).contiguous()
        cfg = _lib.MultiDistCfg(N=N, M=int(M), D=D, k1=float(geom['k1']), sign_convention=int(geom['sign_convention']),
                                unknown_type=UNKNOWN_TYPES[geom['unknown_type']])
        work1 = torch.empty((M, N, N, 2), dtype=torch.float32, device=dev)
        det = torch.empty((D, M, N, N, 2), dtype=torch.float32, device=dev)
        call('adm_multidist_fwd', ctypes.byref(cfg), ptr(obj2d), ptr(probe), ptr(Hp), ptr(work1), ptr(det), stream_ptr())
        ctx.cfg, ctx.geom = cfg, geom
        ctx.save_for_backward(obj2d, probe, Hp, det, dphi)
        return det

    @staticmethod
    def backward(ctx, g_det):
        obj2d, probe, Hp, det, dphi = ctx.saved_tensors
        cfg, geom = ctx.cfg, ctx.geom
        N, M, D = cfg.N, cfg.M, cfg.D
        dev = obj2d.device
        g_det = _f32c(g_det)
        g_obj = g_probe = g_dist = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            Hc = (Hp * torch.tensor([1.0, -1.0], device=dev)).contiguous()
            work1 = torch.empty((M, N, N, 2), dtype=torch.float32, device=dev)
            work2 = torch.empty((D, M, N, N, 2), dtype=torch.float32, device=dev)
            g_obj = torch.empty_like(obj2d) if ctx.needs_input_grad[0] else None
            g_probe = torch.empty_like(probe) if ctx.needs_input_grad[1] else None
            call('adm_multidist_bwd', ctypes.byref(cfg), ptr(obj2d), ptr(probe), ptr(Hc), ptr(g_det), ptr(work1), ptr(work2),
                 ptr(g_obj), ptr(g_probe), stream_ptr())
        if ctx.needs_input_grad[2]:
            # d det_d / d dist = IFFT2(i phi' FFT2(det_d)): the large-frame pass with a vacuum object and H = i phi'
            Hd = torch.stack([torch.zeros_like(dphi), dphi], -1).unsqueeze(0).mul(1.0 / (N * N)).contiguous()
            vac = torch.zeros((N, N, 2), dtype=torch.float32, device=dev)
            if geom['unknown_type'] == 'real_imag':
                vac[..., 0] = 1.0
            cfg1 = _lib.MultiDistCfg(N=N, M=M, D=1, k1=cfg.k1, sign_convention=cfg.sign_convention, unknown_type=cfg.unknown_type)
            work1 = torch.empty((M, N, N, 2), dtype=torch.float32, device=dev)
            dd = torch.empty((1, M, N, N, 2), dtype=torch.float32, device=dev)
            gd = []
            for d in range(D):
                call('adm_multidist_fwd', ctypes.byref(cfg1), ptr(vac), ptr(det[d].contiguous()), ptr(Hd), ptr(work1), ptr(dd),
                     stream_ptr())
                gd.append((g_det[d].double() * dd[0].double()).sum())
            g_dist = (torch.stack(gd) * 1e7).to(torch.float32)              # per cm
        return g_obj, g_probe, g_dist, None


class FrameShiftFn(torch.autograd.Function):
    """out[F,N,N] = Re realign_image_fourier(data[F,N,N], shift[F,2]) (util.py:380-397; forward_model.py:1079-1088)."""

    @staticmethod
    def forward(ctx, data, shift):
        data = _f32c(data)
        shift = _f32c(shift)
        F, N = data.shape[0], data.shape[1]
        if not lf_supported(N) or data.shape[2] != N:
            raise NotImplementedError('adorym_b200: frame shift of %s frames' % (tuple(data.shape),))
        dev = data.device
        order = lf_spectrum_order(N, dev)
        work = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        spec = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        out = torch.empty((F, N, N), dtype=torch.float32, device=dev)
        call('adm_frame_shift_fwd', N, F, ptr(data), ptr(shift), ptr(order), ptr(work), ptr(spec), ptr(out), stream_ptr())
        ctx.save_for_backward(spec, order, shift)
        ctx.dims = (F, N)
        return out

    @staticmethod
    def backward(ctx, g_out):
        spec, order, shift = ctx.saved_tensors
        F, N = ctx.dims
        g_out = _f32c(g_out)
        work = torch.empty((F, N, N, 2), dtype=torch.float32, device=g_out.device)
        g_shift = None
        if ctx.needs_input_grad[1]:
            g_shift = torch.zeros((F, 2), dtype=torch.float32, device=g_out.device)
            call('adm_frame_shift_bwd', N, F, ptr(g_out), ptr(spec), ptr(order), ptr(work), ptr(g_shift), stream_ptr())
        g_data = None
        if ctx.needs_input_grad[0]:
            # the shift of real frames with the real part kept is a real-linear map whose transpose is the shift by -s
            # (the data carry a gradient when an upstream refinable transformation -- the affine alignment -- made them)
            spec2 = torch.empty((F, N, N, 2), dtype=torch.float32, device=g_out.device)
            g_data = torch.empty((F, N, N), dtype=torch.float32, device=g_out.device)
            call('adm_frame_shift_fwd', N, F, ptr(g_out), ptr((-shift).contiguous()), ptr(order), ptr(work), ptr(spec2),
                 ptr(g_data), stream_ptr())
        return g_data, g_shift


class DetectorLossLargeFn(torch.autograd.Function):
    """Large-frame fused |det| + mismatch loss + residuals (adm_detector_loss_large); also differentiable with respect
    to the measured frames (refinable data shift).  Returns (loss, pred[B,N,N], loss_per_pattern[B])."""

    @staticmethod
    def forward(ctx, det, meas, loss_function_type, raw_data_type, poisson_multiplier):
        det = _f32c(det)
        meas = _f32c(meas)
        B, M = det.shape[0], det.shape[1]
        npix = det.shape[2] * det.shape[3]
        inv_n = 1.0 / (B * npix)
        dev = det.device
        pred = torch.empty((B, det.shape[2], det.shape[3]), dtype=torch.float32, device=dev)
        lpp = torch.empty((B,), dtype=torch.float32, device=dev)
        partial = torch.empty((B, 64), dtype=torch.float32, device=dev)
        g_det = torch.empty_like(det)
        g_meas = torch.empty_like(meas) if ctx.needs_input_grad[1] else None
        call('adm_detector_loss_large', B, M, npix, LOSS_TYPES[loss_function_type], RAW_TYPES[raw_data_type],
             float(poisson_multiplier), inv_n, ptr(det), ptr(meas), ptr(pred), ptr(g_det), ptr(g_meas), ptr(partial),
             ptr(lpp), stream_ptr())
        ctx.save_for_backward(g_det, g_meas)
        loss = lpp.sum() * inv_n
        ctx.mark_non_differentiable(pred, lpp)
        return loss, pred, lpp

    @staticmethod
    def backward(ctx, g_loss, g_pred, g_lpp):
        g_det, g_meas = ctx.saved_tensors
        return g_det * g_loss, (g_meas * g_loss if g_meas is not None else None), None, None, None


# ------------------------------------------------------------------------------------------------------
# affine resampling: tilt refinement (wrappers.py:1177-1217) and affine alignment of the data (wrappers.py:1158-1174)
# ------------------------------------------------------------------------------------------------------
class AffineResampleFn(torch.autograd.Function):
    """out = grid_sample(arr, affine_grid(mats), bilinear, border, align_corners=False) on the planes of a dense
    [D0, D1, D2(, 2)] array.  roles = (dim_n, dim_h, dim_w): which memory dimension is the plane index, the rows and the
    columns.  mats: [6] / [2, 3] (shared by all planes) or [N, 2, 3].  Differentiable with respect to arr and mats."""

    @staticmethod
    def forward(ctx, arr, mats, roles):
        arr = _f32c(arr)
        mats = _f32c(mats)
        if arr.dim() == 4 and arr.shape[-1] == 2:
            C = 2
        elif arr.dim() == 3:
            C = 1
        else:
            raise RuntimeError('AffineResampleFn: array must be [D0,D1,D2] or [D0,D1,D2,2], got %s' % (tuple(arr.shape),))
        D = tuple(arr.shape[:3])
        N = D[roles[0]]
        per_plane = int(mats.numel() == 6 * N and mats.dim() == 3)
        if not per_plane and mats.numel() != 6:
            raise RuntimeError('AffineResampleFn: mats must hold one 2x3 matrix, or one per plane')
        out = torch.empty_like(arr)
        call('adm_affine_resample_fwd', D[0], D[1], D[2], C, roles[0], roles[1], roles[2], ptr(arr), ptr(mats), per_plane,
             ptr(out), stream_ptr())
        ctx.geom = (D, C, tuple(roles), per_plane, N, tuple(mats.shape))
        ctx.save_for_backward(arr, mats)
        return out

    @staticmethod
    def backward(ctx, g_out):
        arr, mats = ctx.saved_tensors
        D, C, roles, per_plane, N, mshape = ctx.geom
        g_out = _f32c(g_out)
        g_arr = g_mats = None
        if ctx.needs_input_grad[0]:
            g_arr = torch.zeros_like(arr)
            call('adm_affine_resample_adj', D[0], D[1], D[2], C, roles[0], roles[1], roles[2], ptr(g_out), ptr(mats),
                 per_plane, ptr(g_arr), stream_ptr())
        if ctx.needs_input_grad[1]:
            gm = torch.zeros((N, 6), dtype=torch.float32, device=arr.device)
            call('adm_affine_resample_dmat', D[0], D[1], D[2], C, roles[0], roles[1], roles[2], ptr(arr), ptr(g_out),
                 ptr(mats), per_plane, ptr(gm), stream_ptr())
            g_mats = (gm if per_plane else gm.sum(0)).reshape(mshape)
        return g_arr, g_mats, None


# (plane index, rows, columns) of wrappers.rotate's permutation for an object [Y, X, Z, C], and the sign it gives theta
_TILT_ROLES = {0: ((0, 2, 1), 1.0), 1: ((1, 2, 0), 1.0), 2: ((2, 0, 1), -1.0)}


def rotate_with_grad(obj, theta, axis):
    """adorym/wrappers.py:1177-1217 `rotate` (PyTorch backend): rotation of obj[Y,X,Z,2] about `axis` by the 0-d tensor
    theta in NORMALISED grid coordinates, differentiable with respect to theta (optimize_tilt)."""
    roles, sign = _TILT_ROLES[int(axis)]
    th = theta.reshape(1).to(torch.float32) * sign
    zero = torch.zeros_like(th)
    c, s = torch.cos(th), torch.sin(th)
    mats = torch.cat([c, -s, zero, s, c, zero])
    return AffineResampleFn.apply(obj, mats, roles)


def affine_transform_frames(frames, transform):
    """adorym/wrappers.py:1158-1174 `affine_transform` of a stack of frames [N, H, W] by one [2, 3] matrix."""
    return AffineResampleFn.apply(frames, transform.reshape(6), (0, 1, 2))


# ------------------------------------------------------------------------------------------------------
# ptychography with probes of 128 .. 2048 pixels: slice loop on the batched large-frame kernels
# ------------------------------------------------------------------------------------------------------
def _lf_prepare(h, N, device):
    """[N, N] complex transfer function (natural fftfreq order) -> (Hprep, conj) in the permuted spectrum order, / N^2."""
    order = lf_spectrum_order(N)
    h = np.asarray(h)
    if h.shape != (N, N):
        raise ValueError('transfer function shape %s != (%d, %d)' % (h.shape, N, N))
    hh = h[np.ix_(order, order)]
    hp = torch.from_numpy(np.stack([hh.real, hh.imag], -1).astype('float32')).to(device) * (1.0 / (N * N))
    return hp.contiguous(), (hp * torch.tensor([1.0, -1.0], device=device)).contiguous()


class LargeMultislicePlan(object):
    """MultislicePlan for probe sizes the shared-memory FFT cannot hold (power of two, 128 .. 2048): same attributes, the
    transfer functions prepared for the large-frame kernels."""
    large = True

    def __init__(self, P, n_slices, Y, X, k1, h, device, sign_convention=1, unknown_type='delta_beta', free_prop='none',
                 normalize_fft=False, h_free=None, binning=1, obj_slices=None):
        if not lf_supported(P):
            raise NotImplementedError('adorym_b200: probe size %d is not supported (64, 72, or a power of two from 128 to '
                                      '2048)' % P)
        self.P, self.n_slices, self.Y, self.X = int(P), int(n_slices), int(Y), int(X)
        self.binning = int(binning)
        self.obj_slices = int(obj_slices) if obj_slices is not None else int(n_slices)
        if self.binning < 1 or -(-self.obj_slices // self.binning) != self.n_slices:
            raise ValueError('n_slices must equal ceil(obj_slices / binning)')
        if self.binning > 1 and unknown_type != 'delta_beta':
            raise NotImplementedError('adorym_b200: binning > 1 is accelerated for delta_beta objects only')
        self.k1, self.sign_convention = float(k1), int(sign_convention)
        self.unknown_type = UNKNOWN_TYPES[unknown_type]
        self.free_prop = {'none': 0, 'inf': 1, 'near': 2}[free_prop]
        self.normalize_fft = int(bool(normalize_fft))
        self.device = device
        self.Hp = self.Hc = self.Hfp = self.Hfc = None
        if self.n_slices > 1:
            if h is None:
                raise ValueError('a transfer function is required when n_slices > 1')
            self.Hp, self.Hc = _lf_prepare(h, self.P, device)
        if self.free_prop == 2:
            self.Hfp, self.Hfc = _lf_prepare(h_free, self.P, device)
        # propagate.py:267-270: sign +1 -> fft2 (unnormalised, or 1/N with normalize_fft); sign -1 -> ifft2 (1/N^2, or 1/N),
        # evaluated as conj(fft2(conj(psi))) with the conjugations folded around the same kernels (ff_conj)
        self.ff_conj = self.free_prop == 1 and self.sign_convention == -1
        if self.normalize_fft:
            self.ff_scale = 1.0 / self.P
        else:
            self.ff_scale = 1.0 / (self.P * self.P) if self.sign_convention == -1 else 1.0


class MultisliceLargeFn(torch.autograd.Function):
    """det[B,M,N,N,2] = multislice(obj[Y,X,Z,2], probe[M,N,N,2] | [B,M,N,N,2], pos[B,2] int32) for N = 128 .. 2048: the
    fused object -> transmission volume kernel, then one batched large-frame step per slice (adm_lf_patch_step_fwd), the
    detector transform (adm_lf_farfield_fwd | one vacuum step with the free-space kernel | none).  The wavefield entering
    every slice is kept for the adjoint (adm_lf_patch_step_bwd), whose dL/dT is scatter-added into an un-margined
    gradient volume and taken back to the object by adm_transmission_to_object_adj."""

    @staticmethod
    def forward(ctx, obj, grid, probe, pos, plan):
        obj, probe = _f32c(obj), _f32c(probe)
        if pos.dtype != torch.int32 or not pos.is_cuda:
            raise RuntimeError('pos must be an int32 CUDA tensor [B, 2]')
        pos = pos.contiguous()
        N, Nz, Y, X, Zo = plan.P, plan.n_slices, plan.Y, plan.X, plan.obj_slices
        if tuple(obj.shape) != (Y, X, Zo, 2):
            raise RuntimeError('object must be %s, got %s' % ((Y, X, Zo, 2), tuple(obj.shape)))
        dev = obj.device
        B = pos.shape[0]
        per_pattern = probe.dim() == 5
        M = probe.shape[1] if per_pattern else probe.shape[0]
        if tuple(probe.shape[-3:]) != (N, N, 2) or (per_pattern and probe.shape[0] != B):
            raise RuntimeError('probe must be [M,N,N,2] or [B,M,N,N,2]')
        F = B * M
        tvol = torch.empty((Nz, Y, tvol_pitch(X), 2), dtype=torch.float32, device=dev)
        call('adm_object_to_transmission', Y, X, Zo, ptr(obj), ptr(grid), plan.unknown_type, plan.k1, plan.sign_convention,
             plan.binning, ptr(tvol), stream_ptr())
        psi = (probe if per_pattern else probe.unsqueeze(0).expand(B, M, N, N, 2)).reshape(F, N, N, 2).contiguous()
        work = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        ckpt = []
        for z in range(Nz):
            ckpt.append(psi)
            nxt = torch.empty_like(psi)
            call('adm_lf_patch_step_fwd', N, B, M, Y, X, z, ptr(tvol), ptr(pos), ptr(psi), ptr(plan.Hp) if z < Nz - 1 else None,
                 ptr(work), ptr(nxt), stream_ptr())
            psi = nxt
        vac = None
        if plan.free_prop == 1:
            det = torch.empty_like(psi)
            src = (psi * torch.tensor([1.0, -1.0], device=dev)).contiguous() if plan.ff_conj else psi
            call('adm_lf_farfield_fwd', N, F, ptr(src), ptr(lf_spectrum_order(N, dev)), plan.ff_scale, ptr(work), ptr(det),
                 stream_ptr())
            if plan.ff_conj:
                det = det * torch.tensor([1.0, -1.0], device=dev)
        elif plan.free_prop == 2:
            # free-space propagation of the exit wave (propagate.py:271-280): one step through vacuum (every patch placed
            # outside the object) with the free-space kernel
            vac = torch.full((B, 2), -4 * N - 4 * max(Y, X), dtype=torch.int32, device=dev)
            det = torch.empty_like(psi)
            call('adm_lf_patch_step_fwd', N, B, M, Y, X, 0, ptr(tvol), ptr(vac), ptr(psi), ptr(plan.Hfp), ptr(work), ptr(det),
                 stream_ptr())
        else:
            det = psi
        ctx.plan, ctx.dims, ctx.per_pattern, ctx.has_grid = plan, (B, M), per_pattern, grid is not None
        ctx.probe_shape = tuple(probe.shape)
        saved = [tvol, pos] + ckpt + ([grid] if grid is not None else []) + ([vac] if vac is not None else [])
        ctx.save_for_backward(*saved)
        return det.reshape(B, M, N, N, 2)

    @staticmethod
    def backward(ctx, g_det):
        plan = ctx.plan
        N, Nz, Y, X, Zo = plan.P, plan.n_slices, plan.Y, plan.X, plan.obj_slices
        B, M = ctx.dims
        F = B * M
        saved = list(ctx.saved_tensors)
        tvol, pos = saved[0], saved[1]
        ckpt = saved[2:2 + Nz]
        rest = saved[2 + Nz:]
        grid = rest.pop(0) if ctx.has_grid else None
        vac = rest.pop(0) if plan.free_prop == 2 else None
        dev = tvol.device
        g = _f32c(g_det).reshape(F, N, N, 2)
        work1 = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        work2 = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        if plan.free_prop == 1:
            g2 = torch.empty_like(g)
            # (re, im) -> (re, -im) is its own transpose: the conjugations of the sign -1 convention wrap the adjoint too
            src = (g * torch.tensor([1.0, -1.0], device=dev)).contiguous() if plan.ff_conj else g
            call('adm_lf_farfield_bwd', N, F, ptr(src), ptr(lf_spectrum_order(N, dev)), plan.ff_scale, ptr(work1), ptr(g2),
                 stream_ptr())
            g = g2 * torch.tensor([1.0, -1.0], device=dev) if plan.ff_conj else g2
        elif plan.free_prop == 2:
            g2 = torch.empty_like(g)
            call('adm_lf_patch_step_bwd', N, B, M, Y, X, 0, ptr(tvol), ptr(vac), ptr(ckpt[0]), ptr(plan.Hfc), ptr(g), ptr(work1),
                 ptr(work2), ptr(g2), None, 0, stream_ptr())
            g = g2
        need_obj = ctx.needs_input_grad[0]
        g_tvol = torch.zeros((Nz, Y, gvol_pitch(X, 0), 2), dtype=torch.float32, device=dev) if need_obj else None
        for z in range(Nz - 1, -1, -1):
            g2 = torch.empty_like(g)
            call('adm_lf_patch_step_bwd', N, B, M, Y, X, z, ptr(tvol), ptr(pos), ptr(ckpt[z]),
                 ptr(plan.Hc) if z < Nz - 1 else None, ptr(g), ptr(work1), ptr(work2), ptr(g2), ptr(g_tvol), 0, stream_ptr())
            g = g2
        g_probe = None
        if ctx.needs_input_grad[2]:
            g_probe = g.reshape(B, M, N, N, 2)
            if not ctx.per_pattern:
                g_probe = g_probe.sum(0)
        g_obj = None
        if need_obj:
            g_obj = torch.zeros((Y, X, Zo, 2), dtype=torch.float32, device=dev)
            call('adm_transmission_to_object_adj', Y, X, Zo, 0, ptr(g_tvol), ptr(tvol), ptr(grid), plan.unknown_type, plan.k1,
                 plan.sign_convention, plan.binning, ptr(g_obj), 0, Y, stream_ptr())
        return g_obj, None, g_probe, None, None


class ProbeShiftLargeFn(torch.autograd.Function):
    """ProbeShiftFn for probes of 128 .. 2048 pixels: out[B,M,N,N,2] = Fourier shift of probe[M,N,N,2] by shift[B,2]
    (util.py:380-397) on the large-frame row / column kernels (adm_frame_shift_complex_fwd)."""

    @staticmethod
    def forward(ctx, probe, shift):
        probe, shift = _f32c(probe), _f32c(shift)
        M, N = probe.shape[0], probe.shape[1]
        B = shift.shape[0]
        dev = probe.device
        F = B * M
        frames = probe.unsqueeze(0).expand(B, M, N, N, 2).reshape(F, N, N, 2).contiguous()
        sh = shift.repeat_interleave(M, dim=0).contiguous()
        order = lf_spectrum_order(N, dev)
        work = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        spec = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        out = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        call('adm_frame_shift_complex_fwd', N, F, ptr(frames), ptr(sh), ptr(order), ptr(work), ptr(spec), ptr(out), stream_ptr())
        ctx.save_for_backward(spec, order, sh)
        ctx.dims = (B, M, N)
        return out.reshape(B, M, N, N, 2)

    @staticmethod
    def backward(ctx, g_out):
        spec, order, sh = ctx.saved_tensors
        B, M, N = ctx.dims
        F = B * M
        dev = spec.device
        g = _f32c(g_out).reshape(F, N, N, 2)
        work = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
        g_probe = g_shift = None
        if ctx.needs_input_grad[1]:
            gs = torch.zeros((F, 2), dtype=torch.float32, device=dev)
            call('adm_frame_shift_complex_bwd', N, F, ptr(g), ptr(spec), ptr(order), ptr(work), ptr(gs), stream_ptr())
            g_shift = gs.reshape(B, M, 2).sum(1)
        if ctx.needs_input_grad[0]:
            # the shift is unitary up to its 1/N^2: its adjoint is the shift by -s
            spec2 = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
            gp = torch.empty((F, N, N, 2), dtype=torch.float32, device=dev)
            call('adm_frame_shift_complex_fwd', N, F, ptr(g), ptr((-sh).contiguous()), ptr(order), ptr(work), ptr(spec2), ptr(gp),
                 stream_ptr())
            g_probe = gp.reshape(B, M, N, N, 2).sum(0)
        return g_probe, g_shift


def probe_shift(probe, shift):
    """Per-pattern Fourier shift of the probe modes, dispatching on the probe size."""
    if probe.shape[1] in SUPPORTED_P:
        return ProbeShiftFn.apply(probe, shift)
    return ProbeShiftLargeFn.apply(probe, shift)
