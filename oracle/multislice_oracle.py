"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package `adorym_b200`.

CPU restatement (torch-CPU / numpy, fp32 or fp64) of the reference's multislice forward model,
loss, gradient (reverse-mode AD of the restated forward, as the reference does) and Adam update.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this module, and only as the checker / the reported CPU baseline.

Pinning: the reference (mdw771/adorym @ efbb333b) has NO golden vectors or assertions of its own
(SURVEY.md section 4), so this restatement is pinned against outputs of the reference itself,
imported in the build container through the I/O shims in `oracle/ref_shims`
(`tests/golden/make_golden.py`, `make_golden_multidist.py`, `make_golden_driver.py` -> `tests/golden/*.npz`,
checked by `tests/test_oracle_golden.py` on every CPU run).  Parity against the numpy/HIPS-autograd backend is UNPINNED
(the `autograd` package is not installable here); the PyTorch backend is the pinned one.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import math
import numpy as np
import torch

PI = 3.14159265359  # adorym/constants.py (physics constant used by propagate.py)


# ----------------------------------------------------------------------------------------------
# Fresnel transfer function                                         adorym/propagate.py:54-81
# ----------------------------------------------------------------------------------------------
def gen_freq_mesh(voxel_nm, shape):
    """adorym/propagate.py:54-60."""
    u = np.fft.fftfreq(shape[0])
    v = np.fft.fftfreq(shape[1])
    vv, uu = np.meshgrid(v, u)
    vv = vv / voxel_nm[1]
    uu = uu / voxel_nm[0]
    return uu, vv


def get_kernel(dist_nm, lmbda_nm, voxel_nm, grid_shape, fresnel_approx=True, sign_convention=1):
    """adorym/propagate.py:62-81.  Returns complex128 [P0, P1], unshifted."""
    u, v = gen_freq_mesh(voxel_nm, grid_shape[0:2])
    if fresnel_approx:
        H = np.exp(-sign_convention * 1j * PI * lmbda_nm * dist_nm * (u ** 2 + v ** 2))
    else:
        quad = 1 - lmbda_nm ** 2 * (u ** 2 + v ** 2)
        quad_inner = np.clip(quad, 0, None)
        quad_mask = (quad > 0)
        H = np.exp(sign_convention * 1j * 2 * PI * dist_nm / lmbda_nm * np.sqrt(quad_inner))
        H = H * quad_mask
    return H


# ----------------------------------------------------------------------------------------------
# Rotation lookup table and bilinear resampling                    adorym/util.py:446-552
# ----------------------------------------------------------------------------------------------
def rotation_coords(array_size, theta):
    """LUT of source coordinates for a rotation about axis 0 (y), stored as float16.

    adorym/util.py:446-462 (get_cooridnates_stack_for_rotation, axis=0),
    :465-477 (calculate_original_coordinates_for_rotation, float32 math in torch),
    :514 (cast to float16).  Returns float16 [X*Z, 2]; column 0 = x', column 1 = z'.
    Quirk 19 (util.py:460): the centre subtracted from *both* axes is that of the last axis
    visited by the leaked loop variable -> for axis=0: coordinate along X minus centre[2],
    coordinate along Z minus centre[2]... harmless for X == Z; reproduced literally below.
    """
    Y, X, Z = array_size
    image_center = [(s - 1) / 2 for s in array_size]
    coords_ls = []
    axis = 0
    for this_axis, s in enumerate(array_size):
        if this_axis != axis:
            coord = np.arange(s)
            for i in range(len(array_size)):
                if i != axis and i != this_axis:
                    other_axis = i
                    break
            if other_axis < this_axis:
                coord = np.tile(coord, array_size[other_axis])
            else:
                coord = np.repeat(coord, array_size[other_axis])
            coords_ls.append(coord - image_center[i])
    coord_new = np.stack(coords_ls)  # [2, X*Z] float64
    # wrappers.create_constant / create_variable -> float32 torch tensors (wrappers.py:121-160)
    cn = torch.tensor(coord_new, dtype=torch.float32)
    th = torch.tensor(theta, dtype=torch.float32) if not torch.is_tensor(theta) else theta.float()
    m_rot = torch.tensor([[torch.cos(th), -torch.sin(th)], [torch.sin(th), torch.cos(th)]], dtype=torch.float32)
    coord_old = torch.matmul(m_rot, cn)
    c1 = coord_old[0, :] + image_center[1]
    c2 = coord_old[1, :] + image_center[2]
    out = torch.stack([c1, c2], dim=1).numpy()
    return out.astype('float16')


def grid_normalise(coord_old, X, Z, dtype=torch.float32):
    """The wrapper's LUT -> normalised grid arithmetic, adorym/wrappers.py:1124-1141 with the
    CPU path's float64 upcast of the LUT (util.py:539-540), cast to the array dtype at :1141.

    coord_old float16 [X*Z, 2] (x', z').  The wrapper flips to (z', x') and divides the flipped
    pair by arr.shape[2:] = (X, Z) (sic, wrappers.py:1135-1137); identical when X == Z.
    Returns normalised grid [X*Z, 2] = (gz, gx) in `dtype`.
    """
    g = torch.tensor(np.asarray(coord_old).astype('float64'))
    g = torch.flip(g, (1,))
    arr_shape = torch.tensor([X, Z], dtype=torch.float32)
    g = -1 + 2. * g / arr_shape + 1. / arr_shape
    return g.to(dtype)


def apply_rotation(obj, coord_old):
    """Bilinear, border-clamped resampling of every y-plane of obj[Y,X,Z,C] in the (x,z) plane.

    Restates adorym/util.py:536-552 -> wrappers.grid_sample (wrappers.py:1105-1147) ->
    torch.nn.functional.grid_sample(mode='bilinear', padding_mode='border', align_corners=False)
    as an explicit 4-tap gather (ATen GridSampler: unnormalise ((g+1)*size-1)/2, clip to
    [0,size-1], floor, weights from the se/nw differences, taps outside the image dropped).
    Differentiable w.r.t. obj through index_select.
    """
    Y, X, Z, C = obj.shape
    dt = obj.dtype
    g = grid_normalise(coord_old, X, Z, dtype=dt)
    gz, gx = g[:, 0], g[:, 1]
    iz = ((gz + 1) * Z - 1) / 2
    ix = ((gx + 1) * X - 1) / 2
    iz = torch.clamp(iz, 0, Z - 1)
    ix = torch.clamp(ix, 0, X - 1)
    iz0 = torch.floor(iz)
    ix0 = torch.floor(ix)
    iz1 = iz0 + 1
    ix1 = ix0 + 1
    # ATen naming: "x" is the W axis (= Z here), "y" is the H axis (= X here)
    w_nw = (iz1 - iz) * (ix1 - ix)
    w_ne = (iz - iz0) * (ix1 - ix)
    w_sw = (iz1 - iz) * (ix - ix0)
    w_se = (iz - iz0) * (ix - ix0)
    flat = obj.reshape(Y, X * Z, C)

    def tap(xi, zi, wgt):
        valid = (xi >= 0) & (xi <= X - 1) & (zi >= 0) & (zi <= Z - 1)
        idx = (torch.clamp(xi, 0, X - 1) * Z + torch.clamp(zi, 0, Z - 1)).long()
        wv = (wgt * valid.to(dt)).reshape(1, -1, 1)
        return flat.index_select(1, idx) * wv

    out = tap(ix0, iz0, w_nw) + tap(ix0, iz1, w_ne) + tap(ix1, iz0, w_sw) + tap(ix1, iz1, w_se)
    return out.reshape(Y, X, Z, C)


# ----------------------------------------------------------------------------------------------
# Affine resampling: tilt refinement and affine alignment of the data   adorym/wrappers.py:1158-1217
# ----------------------------------------------------------------------------------------------
def _linspace_base(n, dtype):
    """F.affine_grid(align_corners=False) base coordinates: torch.linspace(-1, 1, n) * (n - 1) / n
    (ATen AffineGridGenerator.cpp linspace_from_neg_one)."""
    return torch.linspace(-1, 1, n, dtype=dtype) * (n - 1) / n


def affine_resample(planes, mats):
    """grid_sample(planes, affine_grid(mats, align_corners=False), bilinear, padding_mode='border',
    align_corners=False) restated as an explicit 4-tap gather.  planes [N, H, W, C]; mats [N, 2, 3] or [2, 3].
    Differentiable w.r.t. planes and mats; as in ATen's clip_coordinates_set_grad the derivative with respect to the
    sampling position is ZERO where the position lies at or beyond the border."""
    N, H, W, C = planes.shape
    dt = planes.dtype
    mats = mats.reshape(-1, 2, 3).to(dt)
    if mats.shape[0] == 1:
        mats = mats.expand(N, 2, 3)
    bx = _linspace_base(W, dt).reshape(1, 1, W)
    by = _linspace_base(H, dt).reshape(1, H, 1)
    m = mats.reshape(N, 6, 1, 1)
    xs = bx * m[:, 0] + by * m[:, 1] + m[:, 2]
    ys = bx * m[:, 3] + by * m[:, 4] + m[:, 5]
    ix = (xs + 1) * (W / 2.) - 0.5           # ATen GridSamplerKernel.cpp (CPU): (in + 1) * (size / 2) - 0.5
    iy = (ys + 1) * (H / 2.) - 0.5

    def clip(v, hi):
        inside = ((v > 0) & (v < hi)).to(dt)
        vc = torch.clamp(v, 0, hi)
        return vc.detach() + (v - v.detach()) * inside

    ix, iy = clip(ix, W - 1), clip(iy, H - 1)
    x0, y0 = torch.floor(ix).detach(), torch.floor(iy).detach()
    x1, y1 = x0 + 1, y0 + 1
    flat = planes.reshape(N, H * W, C)

    def tap(yi, xi, wgt):
        valid = ((xi <= W - 1) & (yi <= H - 1)).to(dt)
        idx = (torch.clamp(yi, 0, H - 1) * W + torch.clamp(xi, 0, W - 1)).long().reshape(N, H * W, 1).expand(N, H * W, C)
        return torch.gather(flat, 1, idx).reshape(N, H, W, C) * (wgt * valid).unsqueeze(-1)

    return (tap(y0, x0, (x1 - ix) * (y1 - iy)) + tap(y0, x1, (ix - x0) * (y1 - iy))
            + tap(y1, x0, (x1 - ix) * (iy - y0)) + tap(y1, x1, (ix - x0) * (iy - y0)))


def rotate_tilt(obj, theta, axis):
    """adorym/wrappers.py:1177-1217 `rotate` (PyTorch backend) on obj[Y, X, Z, C]: the permutation that brings `axis` to
    the front, the sign flip of theta for axis 2 (:1203-1204), m = [[cos, -sin, 0], [sin, cos, 0]] in normalised
    coordinates, affine_grid + bilinear / border grid_sample."""
    perm, sign = {0: ((0, 2, 1, 3), 1.), 1: ((1, 2, 0, 3), 1.), 2: ((2, 0, 1, 3), -1.)}[int(axis)]
    th = torch.as_tensor(theta, dtype=obj.dtype).reshape(()) * sign
    zero = torch.zeros((), dtype=obj.dtype)
    m = torch.stack([torch.stack([torch.cos(th), -torch.sin(th), zero]), torch.stack([torch.sin(th), torch.cos(th), zero])])
    out = affine_resample(obj.permute(*perm), m)
    inv = [perm.index(i) for i in range(4)]
    return out.permute(*inv)


def affine_transform(frames, transform):
    """adorym/wrappers.py:1158-1174 `affine_transform`: frames [N, H, W], one [2, 3] matrix."""
    return affine_resample(frames.unsqueeze(-1), transform)[..., 0]


# ----------------------------------------------------------------------------------------------
# Padding / patch gather                          adorym/util.py:1327-1406, forward_model.py:313-331
# ----------------------------------------------------------------------------------------------
def calculate_pad_len(this_obj_size, probe_pos, probe_size):
    """adorym/util.py:1374-1406 (both unknown_type branches are identical)."""
    probe_pos = np.asarray(probe_pos)
    pad_arr = np.array([[0, 0], [0, 0]])
    if min(probe_pos[:, 0]) < 0:
        pad_arr[0, 0] = -int(min(probe_pos[:, 0]))
    if max(probe_pos[:, 0]) + probe_size[0] > this_obj_size[0]:
        pad_arr[0, 1] = int(max(probe_pos[:, 0])) + probe_size[0] - this_obj_size[0]
    if min(probe_pos[:, 1]) < 0:
        pad_arr[1, 0] = -int(min(probe_pos[:, 1]))
    if max(probe_pos[:, 1]) + probe_size[1] > this_obj_size[1]:
        pad_arr[1, 1] = int(max(probe_pos[:, 1])) + probe_size[1] - this_obj_size[1]
    return pad_arr


def pad_object(obj_rot, this_obj_size, probe_pos, probe_size, unknown_type='delta_beta'):
    """adorym/util.py:1327-1351, PyTorch-backend semantics (quirk 5): vacuum is (0,0) for
    delta_beta and (1,0) for real_imag."""
    pad_arr = calculate_pad_len(this_obj_size, probe_pos, probe_size)
    if np.count_nonzero(pad_arr) > 0:
        p = [0, 0, 0, 0, int(pad_arr[1, 0]), int(pad_arr[1, 1]), int(pad_arr[0, 0]), int(pad_arr[0, 1])]
        if unknown_type == 'delta_beta':
            obj_rot = torch.nn.functional.pad(obj_rot, p, mode='constant', value=0)
        else:
            p3 = p[2:]
            c0 = torch.nn.functional.pad(obj_rot[..., 0], p3, mode='constant', value=1)
            c1 = torch.nn.functional.pad(obj_rot[..., 1], p3, mode='constant', value=0)
            obj_rot = torch.stack([c0, c1], dim=-1)
    return obj_rot, pad_arr


def gather_patches(obj_rot, this_obj_size, pos_batch, probe_size, unknown_type='delta_beta'):
    """adorym/forward_model.py:287,313-331 -> [B, P0, P1, Z, 2]."""
    pos_batch = np.round(np.asarray(pos_batch)).astype(int)
    obj_pad, pad_arr = pad_object(obj_rot, this_obj_size, pos_batch, probe_size, unknown_type)
    subs = []
    for pos in pos_batch:
        py = int(pos[0] + pad_arr[0, 0])
        px = int(pos[1] + pad_arr[1, 0])
        subs.append(obj_pad[py:py + probe_size[0], px:px + probe_size[1], :, :])
    return torch.stack(subs)


# ----------------------------------------------------------------------------------------------
# Fourier shift                                                      adorym/util.py:380-397
# ----------------------------------------------------------------------------------------------
def realign_image_fourier(a, shift, axes=(-2, -1)):
    """adorym/util.py:380-397 on a complex tensor; shift = (y, x)."""
    f = torch.fft.fft2(a, dim=axes)
    s = f.shape
    rdt = a.real.dtype
    fx, fy = np.meshgrid(np.fft.fftfreq(s[axes[1]], 1), np.fft.fftfreq(s[axes[0]], 1))
    fx = torch.tensor(fx, dtype=rdt)
    fy = torch.tensor(fy, dtype=rdt)
    ph = -2 * PI * (fx * shift[1] + fy * shift[0])
    mult = torch.complex(torch.cos(ph), torch.sin(ph))
    return torch.fft.ifft2(f * mult, dim=axes)


# ----------------------------------------------------------------------------------------------
# Multislice                                                         adorym/propagate.py:131-288
# ----------------------------------------------------------------------------------------------
def multislice_propagate_batch(grid_batch, probe, energy_ev, psize_cm, free_prop_cm=None, kernel=None,
                               fresnel_approx=True, binning=1, type='delta_beta', normalize_fft=False,
                               sign_convention=1, scale_ri_by_k=True, shift_exit_wave=None,
                               return_intermediate=False, slice_pos_cm_ls=None, pure_projection=False):
    """adorym/propagate.py:131-288 for pure_projection=False, kappa=None, repeating_slice=None.

    slice_pos_cm_ls given: sparse_multislice_propagate_batch (propagate.py:479-534) -- the propagation after slice i
    uses the paraxial kernel of the gap slice_pos[i+1] - slice_pos[i] (fresnel_propagate_wrapped, :556-568 ->
    get_kernel_wrapped, :84-103), the transmission still uses the voxel size as thickness (:505).

    grid_batch [B, P0, P1, Nz, 2] real; probe complex [P0, P1] or [B, P0, P1]. Returns the
    complex detector-plane wavefield [B, P0, P1] (far field is fftshift-ed, :267-270).
    """
    rdt = grid_batch.dtype
    cdt = torch.complex64 if rdt == torch.float32 else torch.complex128
    grid_shape = grid_batch.shape[1:-1]
    voxel_nm = np.array([psize_cm] * 3) * 1.e7
    lmbda_nm = 1240. / energy_ev
    n_slices = grid_batch.shape[-2]
    delta_nm = voxel_nm[-1]
    size_nm = np.array(grid_shape) * voxel_nm
    if kernel is not None:
        h = kernel
    else:
        h = get_kernel(delta_nm * binning, lmbda_nm, voxel_nm, grid_shape, fresnel_approx=fresnel_approx,
                       sign_convention=sign_convention)
    h = torch.tensor(np.asarray(h)).to(cdt)  # float64 -> array dtype (propagate.py:202-204)
    psi = probe.to(cdt)
    if psi.dim() == 2:
        psi = psi.unsqueeze(0).expand(grid_batch.shape[0], -1, -1)
    inter = []
    if pure_projection:
        # propagate.py:158-186 (delta_beta, not minus-logged, kappa None): one modulation by the line sums, no propagation
        binning = n_slices
    n_steps = int(np.ceil(n_slices / binning))
    for i_step in range(n_steps):
        if return_intermediate:
            inter.append(psi)
        k1 = 2. * PI * delta_nm / lmbda_nm if scale_ri_by_k else 1.
        i_slice = i_step * binning
        this_step = min([binning, n_slices - i_slice])
        d = grid_batch[:, :, :, i_slice:i_slice + this_step, 0]
        b = grid_batch[:, :, :, i_slice:i_slice + this_step, 1]
        if type == 'delta_beta':
            d = d.sum(dim=3)
            b = b.sum(dim=3)
            e = torch.exp(-k1 * b)
            ph = -sign_convention * k1 * d
            c = torch.complex(e * torch.cos(ph), e * torch.sin(ph))
        else:
            c = torch.prod(torch.complex(d, b), dim=3) if this_step > 1 else torch.complex(d[..., 0], b[..., 0])
        psi = psi * c
        if i_step < n_steps - 1 and slice_pos_cm_ls is not None and torch.is_tensor(slice_pos_cm_ls) \
                and slice_pos_cm_ls.requires_grad:
            # refinable slice positions: the gap's paraxial kernel stays in the graph (get_kernel_wrapped, propagate.py:84-103)
            uu, vv = gen_freq_mesh(voxel_nm, grid_shape)
            u2v2 = torch.tensor(uu ** 2 + vv ** 2).to(rdt)
            gap_nm = (slice_pos_cm_ls[i_step + 1] - slice_pos_cm_ls[i_step]) * 1e7
            ph = -sign_convention * PI * lmbda_nm * gap_nm * u2v2
            psi = torch.fft.ifft2(torch.fft.fft2(psi) * torch.complex(torch.cos(ph), torch.sin(ph)))
        elif i_step < n_steps - 1 and slice_pos_cm_ls is not None:
            gap_nm = float(slice_pos_cm_ls[i_step + 1] - slice_pos_cm_ls[i_step]) * 1e7
            hg = get_kernel(gap_nm, lmbda_nm, voxel_nm, grid_shape, sign_convention=sign_convention)
            psi = torch.fft.ifft2(torch.fft.fft2(psi) * torch.tensor(hg).to(cdt))
        elif i_step < n_steps - 1:
            if this_step == binning:
                psi = torch.fft.ifft2(torch.fft.fft2(psi) * h)
            else:
                hh = get_kernel(delta_nm * this_step, lmbda_nm, voxel_nm, grid_shape, sign_convention=sign_convention)
                psi = torch.fft.ifft2(torch.fft.fft2(psi) * torch.tensor(hh).to(cdt))
    if shift_exit_wave is not None:
        psi = realign_image_fourier(psi, shift_exit_wave, axes=(1, 2))
    if free_prop_cm not in [0, None]:
        if isinstance(free_prop_cm, str) and free_prop_cm == 'inf':
            norm = 'ortho' if normalize_fft else None
            if sign_convention == 1:
                psi = torch.fft.fftshift(torch.fft.fft2(psi, norm=norm), dim=(1, 2))
            else:
                psi = torch.fft.fftshift(torch.fft.ifft2(psi, norm=norm), dim=(1, 2))
        else:
            # fresnel_propagate (propagate.py:537-553): always the paraxial kernel (quirk 10)
            dist_nm = float(free_prop_cm) * 1e7
            hf = get_kernel(dist_nm, lmbda_nm, voxel_nm, psi.shape[1:], sign_convention=sign_convention)
            psi = torch.fft.ifft2(torch.fft.fft2(psi) * torch.tensor(hf).to(cdt))
    if return_intermediate:
        return psi, inter
    return psi


# ----------------------------------------------------------------------------------------------
# Forward models                                              adorym/forward_model.py:179-490,819-1037
# ----------------------------------------------------------------------------------------------
def predict_ptycho(obj, probe_real, probe_imag, pos_batch, coord_old, cfg, probe_pos_correction=None,
                   this_i_theta=0, this_ind_batch=None, tilt=None, probe_pos_offset=None, prj_pos_offset=None):
    """PtychographyModel.predict (adorym/forward_model.py:179-387), DP mode, no tilt/defocus refinement.

    cfg keys: this_obj_size, probe_size, energy_ev, psize_cm, h (kernel or None), free_prop_cm,
    unknown_type, binning, normalize_fft, sign_convention, scale_ri_by_k, fresnel_approx,
    optimize_all_probe_pos, two_d_mode.
    Returns magnitudes [B, P0, P1].
    """
    probe_size = cfg['probe_size']
    unknown_type = cfg.get('unknown_type', 'delta_beta')
    pos_batch = np.round(np.asarray(pos_batch)).astype(int)
    if tilt is not None and not cfg.get('two_d_mode', False):
        # optimize_tilt / initial_tilt (forward_model.py:274-276): three w.rotate calls, tilt = tilt_ls[:, this_i_theta]
        obj_rot = rotate_tilt(rotate_tilt(rotate_tilt(obj, tilt[0], 0), tilt[1], 1), tilt[2], 2)
    elif coord_old is not None and not cfg.get('two_d_mode', False):
        obj_rot = apply_rotation(obj, coord_old)
    else:
        obj_rot = obj
    sub = gather_patches(obj_rot, cfg['this_obj_size'], pos_batch, probe_size, unknown_type)
    probe = torch.complex(probe_real, probe_imag)  # [M, P0, P1]
    if probe_pos_offset is not None:
        # optimize_probe_pos_offset (forward_model.py:258-260): one Fourier shift of all modes by the angle's offset
        probe = realign_image_fourier(probe, probe_pos_offset, axes=(1, 2))
    M = probe.shape[0]
    shifted = False
    if probe_pos_correction is not None:
        # quirk 8: trigger is optimize_all_probe_pos or any(correction > 1e-3) (forward_model.py:297,304)
        if cfg.get('optimize_all_probe_pos', False) or bool((probe_pos_correction > 1e-3).any()):
            shifted = True
            pl = []
            for j in range(len(pos_batch)):
                sh = probe_pos_correction[this_i_theta, this_ind_batch[j]]
                pl.append(realign_image_fourier(probe, sh, axes=(1, 2)))
            probe_b = torch.stack(pl)  # [B, M, P0, P1]
    ex_int = None
    for m in range(M):
        p = probe_b[:, m] if shifted else probe[m]
        psi = multislice_propagate_batch(sub, p, cfg['energy_ev'], cfg['psize_cm'], free_prop_cm=cfg.get('free_prop_cm'),
                                         kernel=cfg.get('h'), fresnel_approx=cfg.get('fresnel_approx', True),
                                         binning=cfg.get('binning', 1), type=unknown_type,
                                         normalize_fft=cfg.get('normalize_fft', False),
                                         sign_convention=cfg.get('sign_convention', 1),
                                         scale_ri_by_k=cfg.get('scale_ri_by_k', True),
                                         slice_pos_cm_ls=cfg.get('slice_pos_cm_ls'),
                                         pure_projection=cfg.get('pure_projection', False),
                                         shift_exit_wave=prj_pos_offset)      # optimize_prj_pos_offset (forward_model.py:253-256)
        if M == 1:
            # w.norm -> torch.norm(stack) (wrappers.py:1062-1067): zero subgradient at 0
            return torch.norm(torch.stack([psi.real, psi.imag], dim=0), dim=0)
        i2 = psi.real ** 2 + psi.imag ** 2
        ex_int = i2 if ex_int is None else ex_int + i2
    return torch.sqrt(ex_int)


def predict_fullfield(obj, probe_real, probe_imag, coord_old, cfg):
    """SingleBatchFullfieldModel.predict (adorym/forward_model.py:412-490)."""
    obj_rot = apply_rotation(obj, coord_old) if coord_old is not None else obj
    psi = multislice_propagate_batch(obj_rot.unsqueeze(0), torch.complex(probe_real, probe_imag), cfg['energy_ev'],
                                     cfg['psize_cm'], free_prop_cm=cfg.get('free_prop_cm'), kernel=cfg.get('h'),
                                     fresnel_approx=cfg.get('fresnel_approx', True), binning=cfg.get('binning', 1),
                                     type=cfg.get('unknown_type', 'delta_beta'),
                                     normalize_fft=cfg.get('normalize_fft', False),
                                     sign_convention=cfg.get('sign_convention', 1),
                                     scale_ri_by_k=cfg.get('scale_ri_by_k', True))
    return torch.norm(torch.stack([psi.real, psi.imag], dim=0), dim=0)


def predict_multidist(obj, probe_real, probe_imag, free_prop_cm_ls, cfg):
    """MultiDistModel.predict (adorym/forward_model.py:819-1037) for n_blocks == 1, safe_zone_width == 0,
    forward_algorithm='fresnel', two_d_mode.  Returns [n_dists, P0, P1] magnitudes."""
    probe = torch.complex(probe_real, probe_imag)
    M = probe.shape[0]
    out = []
    refine = cfg.get('optimize_free_prop', False)
    if refine:
        # in-graph paraxial kernel (propagate.py:84-103 get_kernel_wrapped with u_free, v_free of ptychography.py:320-323)
        voxel_nm = np.array([cfg['psize_cm']] * 3) * 1.e7
        uu, vv = gen_freq_mesh(voxel_nm, obj.shape[0:2])
        u2v2 = torch.tensor(uu ** 2 + vv ** 2).to(obj.dtype)
        lmbda_nm = 1240. / cfg['energy_ev']
    for dist in free_prop_cm_ls:
        ex_int = None
        for m in range(M):
            if refine:
                psi = multislice_propagate_batch(obj.unsqueeze(0), probe[m].unsqueeze(0), cfg['energy_ev'], cfg['psize_cm'],
                                                 free_prop_cm=None, kernel=cfg.get('h'), type=cfg.get('unknown_type', 'delta_beta'),
                                                 sign_convention=cfg.get('sign_convention', 1),
                                                 scale_ri_by_k=cfg.get('scale_ri_by_k', True))
                ph = -cfg.get('sign_convention', 1) * PI * lmbda_nm * (dist * 1e7) * u2v2
                psi = torch.fft.ifft2(torch.fft.fft2(psi) * torch.complex(torch.cos(ph), torch.sin(ph)))
                i2 = psi.real ** 2 + psi.imag ** 2
                ex_int = i2 if ex_int is None else ex_int + i2
                continue
            psi = multislice_propagate_batch(obj.unsqueeze(0), probe[m].unsqueeze(0), cfg['energy_ev'], cfg['psize_cm'],
                                             free_prop_cm=float(dist), kernel=cfg.get('h'),
                                             fresnel_approx=cfg.get('fresnel_approx', True), binning=cfg.get('binning', 1),
                                             type=cfg.get('unknown_type', 'delta_beta'),
                                             sign_convention=cfg.get('sign_convention', 1),
                                             scale_ri_by_k=cfg.get('scale_ri_by_k', True))
            i2 = psi.real ** 2 + psi.imag ** 2
            ex_int = i2 if ex_int is None else ex_int + i2
        out.append(torch.sqrt(ex_int))
    return torch.cat(out, 0)


def predict_multidist_tiled(obj, probe_real, probe_imag, free_prop_cm_ls, cfg, pos_batch, subprobe_size, safe_zone_width=0,
                            n_dp_batch=20):
    """MultiDistModel.predict (adorym/forward_model.py:884-1030) for frames sub-divided into tiles (n_blocks > 1) and / or
    a safe zone: the object is padded with vacuum and the probe with (1, 0) (:917-927), every tile reads the region
    [pos - szw, pos + subprobe + szw) of both (:951-963), is propagated on its own, and the safe zone is cropped from the
    magnitudes (:1024-1026).  The reference indexes the real probe patches as [:, i_mode] but the imaginary ones as
    [i_mode] (:1005, quirk 9): with one mode every tile of a sub-batch gets the imaginary patch of the sub-batch's FIRST
    tile -- reproduced here.  One probe mode, sub-batches of at least two tiles (the reference's single-tile branch
    :937-948 cuts an object region of the full probe size and fails on the shape mismatch).
    Returns [n_dists * B, sub_y, sub_x] magnitudes, distance-major."""
    szw = int(safe_zone_width)
    sub = list(subprobe_size)
    tile = [sub[0] + 2 * szw, sub[1] + 2 * szw]
    pos = np.round(np.asarray(pos_batch)).astype(int) - szw
    unknown_type = cfg.get('unknown_type', 'delta_beta')
    Y, X = obj.shape[0], obj.shape[1]
    pad = calculate_pad_len([Y, X], pos, tile)
    obj_p, _ = pad_object(obj, [Y, X, obj.shape[2]], pos, tile, unknown_type)
    pw = [int(pad[1][0]), int(pad[1][1]), int(pad[0][0]), int(pad[0][1])]
    pr_p = torch.nn.functional.pad(probe_real, pw, value=1.)
    pi_p = torch.nn.functional.pad(probe_imag, pw, value=0.)
    assert probe_real.shape[0] == 1, 'quirk 9: the reference only runs with one probe mode here'
    chunks = [list(range(i, min(i + n_dp_batch, len(pos)))) for i in range(0, len(pos), n_dp_batch)]
    out = []
    for dist in free_prop_cm_ls:
        for ch in chunks:
            so, sr, si = [], [], []
            for j in ch:
                y0, x0 = pos[j][0] + int(pad[0][0]), pos[j][1] + int(pad[1][0])
                so.append(obj_p[y0:y0 + tile[0], x0:x0 + tile[1]])
                sr.append(pr_p[0, y0:y0 + tile[0], x0:x0 + tile[1]])
                si.append(pi_p[0, y0:y0 + tile[0], x0:x0 + tile[1]])
            so, sr = torch.stack(so), torch.stack(sr)
            si = si[0].unsqueeze(0).expand_as(sr)            # quirk 9
            psi = multislice_propagate_batch(so, torch.complex(sr, si), cfg['energy_ev'], cfg['psize_cm'],
                                             free_prop_cm=float(dist), kernel=cfg.get('h'),
                                             type=unknown_type, sign_convention=cfg.get('sign_convention', 1),
                                             scale_ri_by_k=cfg.get('scale_ri_by_k', True))
            out.append(torch.sqrt(psi.real ** 2 + psi.imag ** 2))
    out = torch.cat(out, 0)
    if szw > 0:
        out = out[:, szw:szw + sub[0], szw:szw + sub[1]]
    return out


# ----------------------------------------------------------------------------------------------
# Loss and regularisers               adorym/forward_model.py:88-147, regularizers.py:30-110
# ----------------------------------------------------------------------------------------------
def mismatch_loss(pred, prj_abs, loss_function_type='lsq', raw_data_type='magnitude', poisson_multiplier=1.):
    """adorym/forward_model.py:88-103; prj_abs is already abs(prj) (forward_model.py:116)."""
    if loss_function_type == 'lsq':
        if raw_data_type == 'magnitude':
            return torch.mean((pred - prj_abs) ** 2)
        return torch.mean((pred - torch.sqrt(prj_abs)) ** 2)
    pm = poisson_multiplier
    if raw_data_type == 'magnitude':
        return torch.mean(pred ** 2 * pm - prj_abs ** 2 * pm * torch.log(pred ** 2 * pm))
    return torch.mean(pred ** 2 * pm - prj_abs * pm * torch.log(pred ** 2 * pm))


def total_variation_3d(arr):
    """adorym/util.py:1427-1440."""
    res = torch.sum(torch.abs(torch.roll(arr, 1, dims=0) - arr))
    res = res + torch.sum(torch.abs(torch.roll(arr, 1, dims=1) - arr))
    res = res + torch.sum(torch.abs(torch.roll(arr, 1, dims=2) - arr))
    return res / arr.numel()


def regularization_value(obj, alpha_d=None, alpha_b=None, gamma=None, weight_l1=None, unknown_type='delta_beta'):
    """L1 (regularizers.py:30-46), reweighted L1 (:64-83) and TV (:95-110), delta_beta type."""
    assert unknown_type == 'delta_beta'
    reg = torch.zeros((), dtype=obj.dtype)
    w0 = 1. if weight_l1 is None else weight_l1[..., 0]
    w1 = 1. if weight_l1 is None else weight_l1[..., 1]
    if alpha_d not in [None, 0]:
        reg = reg + alpha_d * torch.mean(w0 * torch.abs(obj[..., 0]))
    if alpha_b not in [None, 0]:
        reg = reg + alpha_b * torch.mean(w1 * torch.abs(obj[..., 1]))
    if gamma not in [None, 0]:
        reg = reg + gamma * total_variation_3d(obj[..., 0])
        reg = reg + gamma * total_variation_3d(obj[..., 1])
    return reg


def multidist_loss_and_grads(obj, probe_real, probe_imag, free_prop_cm_ls, prj_abs, cfg, probe_pos_correction=None,
                             want=('obj',), prj_affine_ls=None, tiles=None):
    """MultiDistModel.get_loss_function (adorym/forward_model.py:1039-1092) for n_blocks == 1, no affine / offset
    refinement: predict (:819-1037), optional per-distance Fourier shift of the measured frames with the real part
    kept (:1079-1088, optimize_all_probe_pos; probe_pos_correction is [n_dists, 2]), mismatch loss (:88-103 applies
    abs() to the -- possibly shifted -- data).  prj_abs [n_dists, P0, P1].  Returns loss, pred, grads."""
    leaves = {'obj': obj.clone().requires_grad_('obj' in want),
              'probe_real': probe_real.clone().requires_grad_('probe_real' in want),
              'probe_imag': probe_imag.clone().requires_grad_('probe_imag' in want)}
    if 'free_prop_cm' in want:
        leaves['free_prop_cm'] = torch.as_tensor(np.asarray(free_prop_cm_ls, dtype='float64')).to(obj.dtype).requires_grad_(True)
        free_prop_cm_ls = leaves['free_prop_cm']
        cfg = dict(cfg, optimize_free_prop=True)
    if tiles is not None:
        # tiles = dict(pos_batch, subprobe_size, safe_zone_width, n_dp_batch): sub-divided frames / safe zone
        pred = predict_multidist_tiled(leaves['obj'], leaves['probe_real'], leaves['probe_imag'], free_prop_cm_ls, cfg, **tiles)
    else:
        pred = predict_multidist(leaves['obj'], leaves['probe_real'], leaves['probe_imag'], free_prop_cm_ls, cfg)
    prj = prj_abs
    if prj_affine_ls is not None:
        # optimize_prj_affine (forward_model.py:1066-1072): every distance's frame resampled by its own 2x3 matrix
        leaves['prj_affine_ls'] = prj_affine_ls.clone().requires_grad_('prj_affine_ls' in want)
        prj = torch.cat([affine_transform(prj[i:i + 1], leaves['prj_affine_ls'][i]) for i in range(len(free_prop_cm_ls))])
    if probe_pos_correction is not None:
        leaves['probe_pos_correction'] = probe_pos_correction.clone().requires_grad_('probe_pos_correction' in want)
        out = []
        for i in range(len(free_prop_cm_ls)):
            fr = prj[i:i + 1]
            sh = realign_image_fourier(torch.complex(fr, torch.zeros_like(fr)), leaves['probe_pos_correction'][i],
                                       axes=(1, 2))
            out.append(sh.real)
        prj = torch.cat(out)
    loss = mismatch_loss(pred, torch.abs(prj), cfg.get('loss_function_type', 'lsq'), cfg.get('raw_data_type', 'magnitude'),
                         cfg.get('poisson_multiplier', 1.))
    grads = torch.autograd.grad(loss, [leaves[k] for k in want], allow_unused=True)
    return loss.detach(), pred.detach(), dict(zip(want, grads))


# ----------------------------------------------------------------------------------------------
# Adam                                                               adorym/optimizers.py:309-318
# ----------------------------------------------------------------------------------------------
def adam_step(x, g, m, v, i_batch, step_size=0.001, b1=0.9, b2=0.999, eps=1e-7):
    """adorym/optimizers.py:309-318.  Returns (x_new, m_new, v_new)."""
    m = b1 * m
    m = m + (1 - b1) * g
    v = b2 * v
    v = v + (1 - b2) * (g ** 2)
    mhat = m / (1 - b1 ** (i_batch + 1))
    vhat = v / (1 - b2 ** (i_batch + 1))
    d = step_size * mhat / (torch.sqrt(vhat) + eps)
    return x - d, m, v


# ----------------------------------------------------------------------------------------------
# Whole-loss + gradients (what Differentiator.get_gradients returns, wrappers.py:300-331)
# ----------------------------------------------------------------------------------------------
def ptycho_loss_and_grads(obj, probe_real, probe_imag, pos_batch, prj_abs, coord_old, cfg, reg=None,
                          probe_pos_correction=None, this_i_theta=0, this_ind_batch=None,
                          want=('obj', 'probe_real', 'probe_imag'), tilt=None, probe_pos_offset=None, prj_pos_offset=None):
    """Loss of calculate_loss (forward_model.py:389-401) and its gradient w.r.t. `want`.  tilt: [3] angles of this
    rotation angle (tilt_ls[:, i_theta]); probe_pos_offset: [2] offset of this angle (probe_pos_offset[i_theta])."""
    leaves = {'obj': obj.clone().requires_grad_('obj' in want),
              'probe_real': probe_real.clone().requires_grad_('probe_real' in want),
              'probe_imag': probe_imag.clone().requires_grad_('probe_imag' in want)}
    ppc = None
    if probe_pos_correction is not None:
        ppc = probe_pos_correction.clone().requires_grad_('probe_pos_correction' in want)
        leaves['probe_pos_correction'] = ppc
    if tilt is not None:
        tilt = leaves['tilt'] = torch.as_tensor(tilt).to(obj.dtype).clone().requires_grad_('tilt' in want)
    if probe_pos_offset is not None:
        probe_pos_offset = leaves['probe_pos_offset'] = torch.as_tensor(probe_pos_offset).to(obj.dtype).clone().requires_grad_(
            'probe_pos_offset' in want)
    if 'slice_pos_cm_ls' in want:
        leaves['slice_pos_cm_ls'] = torch.as_tensor(np.asarray(cfg['slice_pos_cm_ls'], dtype='float64')).to(obj.dtype).clone().requires_grad_(True)
        cfg = dict(cfg, slice_pos_cm_ls=leaves['slice_pos_cm_ls'])
    if prj_pos_offset is not None:
        prj_pos_offset = leaves['prj_pos_offset'] = torch.as_tensor(prj_pos_offset).to(obj.dtype).clone().requires_grad_(
            'prj_pos_offset' in want)
    pred = predict_ptycho(leaves['obj'], leaves['probe_real'], leaves['probe_imag'], pos_batch, coord_old, cfg,
                          probe_pos_correction=ppc, this_i_theta=this_i_theta, this_ind_batch=this_ind_batch, tilt=tilt,
                          probe_pos_offset=probe_pos_offset, prj_pos_offset=prj_pos_offset)
    pred_l, prj_l = pred, prj_abs
    if cfg.get('beamstop', None) is not None:
        # forward_model.py:128-136: pixels with beamstop < 1e-5 are dropped from the mean
        bs = torch.as_tensor(np.asarray(cfg['beamstop']) >= 1e-5)
        sel = bs.unsqueeze(0).expand_as(pred)
        pred_l = pred[sel].reshape(pred.shape[0], -1)
        prj_l = prj_abs[sel].reshape(pred.shape[0], -1)
    loss = mismatch_loss(pred_l, prj_l, cfg.get('loss_function_type', 'lsq'), cfg.get('raw_data_type', 'magnitude'),
                         cfg.get('poisson_multiplier', 1.))
    data_loss = loss.detach().clone()
    if reg:
        loss = loss + regularization_value(leaves['obj'], **reg)
    grads = torch.autograd.grad(loss, [leaves[k] for k in want], allow_unused=True)
    return loss.detach(), data_loss, pred.detach(), dict(zip(want, grads))


# ----------------------------------------------------------------------------------------------
# Minibatch schedule                                   adorym/ptychography.py:786-850,901-909
# ----------------------------------------------------------------------------------------------
def epoch_schedule(i_epoch, n_theta, n_pos, minibatch_size, n_ranks, two_d_mode=False, theta_ls=None,
                   randomize_probe_pos=False, update_scheme='immediate'):
    """Returns the list of global chunks (each [<=minibatch*n_ranks, 2] int32 of (i_theta, i_spot)).

    adorym/ptychography.py:786-847.  The numpy global RNG is seeded with i_epoch (:791), angles are
    shuffled (:795), each angle's spot list is padded with random re-draws to a multiple of the
    minibatch (:816-823), and the concatenation is split into chunks of minibatch*n_ranks (:846).
    """
    n_tot = minibatch_size * n_ranks
    rs = np.random.RandomState(i_epoch)  # np.random.seed(i_epoch) + legacy global functions
    if not two_d_mode:
        theta_ind_ls = np.arange(n_theta)
        rs.shuffle(theta_ind_ls)
    else:
        theta_ind_ls = np.array([0])
    rows = []
    for i_theta in theta_ind_ls:
        spots = np.arange(n_pos)
        if randomize_probe_pos:
            spots = rs.choice(spots, len(spots), replace=False)
        if update_scheme == 'immediate' and n_pos % minibatch_size != 0:
            spots = np.append(spots, rs.choice(spots[:-n_pos % minibatch_size],
                                               minibatch_size - (n_pos % minibatch_size), replace=False))
        elif update_scheme == 'per angle' and n_pos % n_tot != 0:
            spots = np.append(spots, rs.choice(spots[:-n_pos % n_tot], n_tot - (n_pos % n_tot), replace=False))
        rows.append(np.stack([np.array([i_theta] * len(spots)), spots], axis=1))
    ind = np.concatenate(rows, axis=0).astype('int32')
    chunks = [ind[i:i + n_tot] for i in range(0, len(ind), n_tot)]
    return chunks


def rank_slice_of_chunk(chunks, i_batch, rank, minibatch_size, n_ranks):
    """adorym/ptychography.py:901-909: top up a short chunk from chunk 0, pick this rank's entries,
    sort the spot indices."""
    n_tot = minibatch_size * n_ranks
    c = chunks[i_batch]
    if len(c) < n_tot:
        c = np.concatenate([c, chunks[0][:n_tot - len(c)]])
    i_theta = int(c[rank * minibatch_size, 0])
    ind = np.sort(c[rank * minibatch_size:(rank + 1) * minibatch_size, 1])
    return i_theta, ind


# ----------------------------------------------------------------------------------------------
# Probe / object initialisation helpers                    adorym/util.py:189-206, :71-124
# ----------------------------------------------------------------------------------------------
def gaussian_probe(size, mag_max, mag_sigma, phase_max, phase_sigma):
    """adorym/util.py:189-195 + mag_phase_to_real_imag."""
    py = np.arange(size[0]) - (size[0] - 1.) / 2
    px = np.arange(size[1]) - (size[1] - 1.) / 2
    pxx, pyy = np.meshgrid(px, py)
    mag = mag_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * mag_sigma ** 2))
    ph = phase_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * phase_sigma ** 2))
    return mag * np.cos(ph), mag * np.sin(ph)
