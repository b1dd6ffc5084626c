"""Host-side mirror of adorym/propagate.py for the hot path: same function names, argument meaning and error
behaviour; the arithmetic runs in the sm_100a kernels (adorym_b200/csrc/multislice.cu)."""
import numpy as np
import torch

from . import ops

PI = ops.PI


def gen_freq_mesh(voxel_nm, shape):
    """adorym/propagate.py:54-60."""
    u = np.fft.fftfreq(shape[0])
    v = np.fft.fftfreq(shape[1])
    vv, uu = np.meshgrid(v, u)
    return uu / voxel_nm[0], vv / voxel_nm[1]


def get_kernel(dist_nm, lmbda_nm, voxel_nm, grid_shape, fresnel_approx=True, sign_convention=1):
    """Unshifted transfer function, complex128 [P0, P1] (adorym/propagate.py:62-81)."""
    u, v = gen_freq_mesh(voxel_nm, grid_shape[0:2])
    r2 = u ** 2 + v ** 2
    if fresnel_approx:
        return np.exp(-sign_convention * 1j * PI * lmbda_nm * dist_nm * r2)
    quad = 1 - lmbda_nm ** 2 * r2
    H = np.exp(sign_convention * 1j * 2 * PI * dist_nm / lmbda_nm * np.sqrt(np.clip(quad, 0, None)))
    return H * (quad > 0)


def _free_prop_mode(free_prop_cm):
    if free_prop_cm in [0, None]:
        return 'none'
    if isinstance(free_prop_cm, str):
        if free_prop_cm == 'inf':
            return 'inf'
        raise ValueError('free_prop_cm must be a number, None or "inf"')
    return 'near'


_plan_cache = {}


def _kernel_digest(kernel):
    """Content hash of a user-supplied transfer function (an id() can be recycled by the allocator and would then hit
    a stale plan)."""
    if kernel is None:
        return None
    import hashlib
    return hashlib.sha1(np.ascontiguousarray(np.asarray(kernel)).tobytes()).hexdigest()


def multislice_propagate_batch(grid_batch, probe_real, probe_imag, energy_ev, psize_cm, delta_cm=None,
                               free_prop_cm=None, obj_batch_shape=None, kernel=None, fresnel_approx=True,
                               pure_projection=False, binning=1, device=None, type='delta_beta',
                               normalize_fft=False, sign_convention=1, optimize_free_prop=False, u_free=None, v_free=None,
                               scale_ri_by_k=True, is_minus_logged=False, pure_projection_return_sqrt=False,
                               kappa=None, repeating_slice=None, return_fft_time=False, shift_exit_wave=None,
                               return_intermediate_wavefields=False):
    """Drop-in for adorym/propagate.py:131-288 (same signature, returns [real, imag] each [B, P, P]).

    grid_batch is the already-gathered patch stack [B, P, P, Nz, 2]; it is viewed as a slice-major volume with the
    B patches stacked along y, so the same fused kernel serves (the model classes skip this repacking and hand the
    rotated volume + probe positions to the kernel directly).  Options outside the hot path raise
    NotImplementedError naming the option (no silent fallback).
    """
    for name, val, ok in (('pure_projection', pure_projection, False), ('kappa', kappa, None),
                          ('repeating_slice', repeating_slice, None), ('shift_exit_wave', shift_exit_wave, None),
                          ('optimize_free_prop', optimize_free_prop, False), ('delta_cm', delta_cm, None),
                          ('return_intermediate_wavefields', return_intermediate_wavefields, False)):
        if val is not ok and val != ok:
            raise NotImplementedError('adorym_b200.multislice_propagate_batch: option %s=%r is not on the accelerated '
                                      'path' % (name, val))
    if binning != 1:
        raise NotImplementedError('adorym_b200.multislice_propagate_batch: binning > 1 is not on the accelerated path')
    if type not in ('delta_beta', 'real_imag'):
        raise ValueError('unknown_type must be delta_beta or real_imag.')
    if not grid_batch.is_cuda:
        raise RuntimeError('adorym_b200 has no CPU path: grid_batch must be a CUDA tensor')
    B, P0, P1, Nz, _ = grid_batch.shape
    if P0 != P1:
        raise NotImplementedError('adorym_b200: only square probes are supported')
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    lmbda_nm = 1240. / energy_ev
    mode = _free_prop_mode(free_prop_cm)
    h = kernel if kernel is not None else get_kernel(voxel_nm[-1] * binning, lmbda_nm, voxel_nm, [P0, P1],
                                                    fresnel_approx=fresnel_approx, sign_convention=sign_convention)
    h_free = None
    if mode == 'near':
        # fresnel_propagate ignores fresnel_approx (quirk 10, propagate.py:546)
        h_free = get_kernel(float(free_prop_cm) * 1e7, lmbda_nm, voxel_nm, [P0, P1], sign_convention=sign_convention)
    key = (P0, Nz, B, float(energy_ev), float(psize_cm), mode, float(free_prop_cm) if mode == 'near' else 0., type,
           bool(normalize_fft), int(sign_convention), bool(scale_ri_by_k), bool(fresnel_approx),
           _kernel_digest(kernel), str(grid_batch.device))
    large = P0 not in ops.SUPPORTED_P          # 128 .. 2048 (power of two): the batched large-frame slice steps
    plan = _plan_cache.get(key)
    if plan is None:
        Plan = ops.LargeMultislicePlan if large else ops.MultislicePlan
        plan = Plan(P0, Nz, B * P0, P1, ops.k1_of(energy_ev, psize_cm, scale_ri_by_k), h, grid_batch.device,
                    sign_convention=sign_convention, unknown_type=type, free_prop=mode, normalize_fft=normalize_fft,
                    h_free=h_free)
        if len(_plan_cache) >= 32:          # bounded: drop the oldest plan
            _plan_cache.pop(next(iter(_plan_cache)))
        _plan_cache[key] = plan
    pos = torch.stack([torch.arange(B, dtype=torch.int32) * P0, torch.zeros(B, dtype=torch.int32)], 1).to(grid_batch.device)
    if probe_real.dim() == 2:
        probe = torch.stack([probe_real, probe_imag], -1).unsqueeze(0)          # [1, P, P, 2]
    else:
        probe = torch.stack([probe_real, probe_imag], -1).unsqueeze(1)          # [B, 1, P, P, 2]
    if large:
        # the patch stack [B, P, P, Nz, 2] IS an object of B * P rows in the reference layout [Y][X][Z][2]
        det = ops.MultisliceLargeFn.apply(grid_batch.reshape(B * P0, P1, Nz, 2), None, probe, pos, plan)
    else:
        rot = grid_batch.permute(3, 0, 1, 2, 4).reshape(Nz, B * P0, P1, 2).contiguous()
        det = ops.MultisliceFn.apply(rot, None, probe, pos, plan, None, 'zyx')
    ret = [det[:, 0, :, :, 0], det[:, 0, :, :, 1]]
    if return_fft_time:
        ret.append(0.)
    return ret
