"""Host-side mirror of the adorym/util.py helpers that sit on the hot path (same names and argument meaning)."""
import os
import numpy as np
import torch

from . import ops

PI = ops.PI


# ------------------------------------------------------------------------------------------------------
# rotation lookup tables                                                     adorym/util.py:446-525
# ------------------------------------------------------------------------------------------------------
def get_cooridnates_stack_for_rotation(array_size, axis=0):
    """Centred (x, z) index stack [2, X*Z] for a rotation about `axis` = 0 (adorym/util.py:446-462, sic).

    Quirk 19 is reproduced: each in-plane coordinate is centred with the centre of the *other* in-plane axis
    (harmless when X == Z)."""
    if axis != 0:
        raise NotImplementedError('adorym_b200: only rotation about axis 0 is on the accelerated path')
    Y, X, Z = array_size
    cx, cz = (X - 1) / 2, (Z - 1) / 2
    c1 = np.repeat(np.arange(X), Z) - cz
    c2 = np.tile(np.arange(Z), X) - cx
    return np.stack([c1, c2])


def calculate_original_coordinates_for_rotation(array_size, coord_new, theta):
    """fp32 torch-CPU arithmetic as in adorym/util.py:465-477 -> ndarray [X*Z, 2] float32 (x', z')."""
    cx, cz = (array_size[1] - 1) / 2, (array_size[2] - 1) / 2
    th = torch.as_tensor(theta, dtype=torch.float32)
    m_rot = torch.tensor([[torch.cos(th), -torch.sin(th)], [torch.sin(th), torch.cos(th)]], dtype=torch.float32)
    cn = torch.as_tensor(np.asarray(coord_new), dtype=torch.float32)
    co = torch.matmul(m_rot, cn)
    return torch.stack([co[0] + cx, co[1] + cz], dim=1).numpy()


def rotation_lookup(array_size, theta, reverse=False):
    """One LUT as the reference stores it: float16 [X*Z, 2] (adorym/util.py:507-515)."""
    coord_new = get_cooridnates_stack_for_rotation(array_size, axis=0)
    th = np.float32(theta)
    return calculate_original_coordinates_for_rotation(array_size, coord_new, -th if reverse else th).astype('float16')


def save_rotation_lookup(array_size, theta_ls, dest_folder=None, override_backend=None):
    """adorym/util.py:492-516: writes {theta:.5f}.npy and _{theta:.5f}.npy (inverse angle) as float16."""
    n_theta = len(theta_ls)
    if dest_folder is None:
        dest_folder = 'arrsize_{}_{}_{}_ntheta_{}'.format(array_size[0], array_size[1], array_size[2], n_theta)
    if not os.path.exists(dest_folder):
        os.mkdir(dest_folder)
    for theta in np.asarray(theta_ls, dtype='float32'):
        np.save(os.path.join(dest_folder, '{:.5f}'.format(theta)), rotation_lookup(array_size, theta))
        np.save(os.path.join(dest_folder, '_{:.5f}'.format(theta)), rotation_lookup(array_size, theta, reverse=True))
    return None


def read_origin_coords(src_folder, theta, reverse=False):
    """adorym/util.py:519-525."""
    name = '{:.5f}.npy'.format(theta) if not reverse else '_{:.5f}.npy'.format(theta)
    return np.load(os.path.join(src_folder, name), allow_pickle=True)


class RotationTable(object):
    """Device-resident cache of normalised rotation grids, one per angle (replaces the per-`predict` np.load of
    adorym/forward_model.py:233-235).  Values are identical to what the reference derives from its fp16 files."""

    def __init__(self, array_size, theta_ls, device, src_folder=None):
        self.array_size = list(array_size)
        self.theta_ls = np.asarray(theta_ls)
        self.device = device
        self.src_folder = src_folder
        self._cache = {}

    def lut(self, i_theta, reverse=False):
        theta = self.theta_ls[i_theta]
        if self.src_folder is not None and os.path.isdir(self.src_folder):
            try:
                return read_origin_coords(self.src_folder, theta, reverse=reverse)
            except (IOError, OSError):
                pass
        return rotation_lookup(self.array_size, theta, reverse=reverse)

    def grid(self, i_theta):
        g = self._cache.get(int(i_theta))
        if g is None:
            g = ops.normalized_grid(self.lut(i_theta), self.array_size[1], self.array_size[2], self.device)
            self._cache[int(i_theta)] = g
        return g


def apply_rotation(obj, coord_old, interpolation='bilinear', axis=0, device=None, reverse=False, override_backend=None):
    """Drop-in for adorym/util.py:536-552: obj [Y,X,Z,2] -> rotated [Y,X,Z,2] (reference layout).

    `reverse=True` applies the exact transpose (util.py:623-696)."""
    if interpolation != 'bilinear' or axis != 0:
        raise NotImplementedError('adorym_b200.apply_rotation: only bilinear rotation about axis 0 is accelerated')
    if not obj.is_cuda:
        raise RuntimeError('adorym_b200 has no CPU path: obj must be a CUDA tensor')
    Y, X, Z, _ = obj.shape
    grid = ops.normalized_grid(coord_old, X, Z, obj.device)
    if not reverse:
        return ops.RotateFn.apply(obj, grid).permute(1, 2, 0, 3).contiguous()
    g_obj = torch.zeros((Y, X, Z, 2), dtype=torch.float32, device=obj.device)
    src = obj.permute(2, 0, 1, 3).contiguous()
    ops.call('adm_rotate_adj', Y, X, Z, ops.ptr(src), ops.ptr(grid), ops.ptr(g_obj), ops.stream_ptr())
    return g_obj


# ------------------------------------------------------------------------------------------------------
# padding                                                                    adorym/util.py:1327-1406
# ------------------------------------------------------------------------------------------------------
def calculate_pad_len(this_obj_size, probe_pos, probe_size, unknown_type='delta_beta'):
    """adorym/util.py:1374-1406."""
    probe_pos = np.asarray(probe_pos)
    pad_arr = np.array([[0, 0], [0, 0]])
    for ax in (0, 1):
        lo = min(probe_pos[:, ax])
        hi = max(probe_pos[:, ax])
        if lo < 0:
            pad_arr[ax, 0] = -int(lo)
        if hi + probe_size[ax] > this_obj_size[ax]:
            pad_arr[ax, 1] = int(hi) + probe_size[ax] - this_obj_size[ax]
    return pad_arr


def pad_object(obj_rot, this_obj_size, probe_pos, probe_size, mode='constant', unknown_type='delta_beta',
               override_backend=None):
    """Drop-in for adorym/util.py:1327-1351 (PyTorch-backend semantics: vacuum is (0,0) / (1,0)).
    The fused multislice kernel never materialises this padded copy (it predicates the loads instead); this
    function exists for API compatibility."""
    if mode != 'constant':
        raise NotImplementedError('adorym_b200.pad_object: mode=%r' % mode)
    pad_arr = calculate_pad_len(this_obj_size, probe_pos, probe_size, unknown_type)
    if np.count_nonzero(pad_arr) > 0:
        p = [0, 0] * (obj_rot.dim() - 2) + [int(pad_arr[1, 0]), int(pad_arr[1, 1]), int(pad_arr[0, 0]), int(pad_arr[0, 1])]
        if unknown_type == 'delta_beta':
            obj_rot = torch.nn.functional.pad(obj_rot, p, mode='constant', value=0)
        elif unknown_type == 'real_imag':
            c0 = torch.nn.functional.pad(obj_rot[..., 0], p[2:], mode='constant', value=1)
            c1 = torch.nn.functional.pad(obj_rot[..., 1], p[2:], mode='constant', value=0)
            obj_rot = torch.stack([c0, c1], dim=-1)
    return obj_rot, pad_arr


# ------------------------------------------------------------------------------------------------------
# initial guesses                                                            adorym/util.py:71-124,189-283
# ------------------------------------------------------------------------------------------------------
def mag_phase_to_real_imag(mag, phase):
    return mag * np.cos(phase), mag * np.sin(phase)


def generate_gaussian_map(size, mag_max, mag_sigma, phase_max, phase_sigma):
    """adorym/util.py:189-195."""
    py = np.arange(size[0]) - (size[0] - 1.) / 2
    px = np.arange(size[1]) - (size[1] - 1.) / 2
    pxx, pyy = np.meshgrid(px, py)
    map_mag = mag_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * mag_sigma ** 2))
    map_phase = phase_max * np.exp(-(pxx ** 2 + pyy ** 2) / (2 * phase_sigma ** 2))
    return map_mag, map_phase


def generate_disk(shape, radius):
    yy, xx = np.mgrid[:shape[0], :shape[1]]
    cy, cx = (np.array(shape) - 1.) / 2
    return ((yy - cy) ** 2 + (xx - cx) ** 2 <= radius ** 2).astype('float64')


def initialize_object_for_dp(this_obj_size, object_type='normal', initial_guess=None,
                             random_guess_means_sigmas=(8.7e-7, 5.1e-8, 1e-7, 1e-8), unknown_type='delta_beta',
                             non_negativity=False, seed=None):
    """adorym/util.py:71-124 (first level only).  The reference seeds numpy with int(time()/60) (quirk 6);
    pass `seed` for reproducibility or `initial_guess` for parity runs."""
    if initial_guess is None:
        import time
        rs = np.random.RandomState(int(time.time() / 60) if seed is None else seed)
        m = random_guess_means_sigmas
        obj_delta = rs.normal(size=this_obj_size, loc=m[0], scale=m[2])
        obj_beta = rs.normal(size=this_obj_size, loc=m[1], scale=m[3])
    else:
        obj_delta = np.array(initial_guess[0])
        obj_beta = np.array(initial_guess[1])
    if object_type == 'phase_only':
        if unknown_type == 'delta_beta':
            obj_beta[...] = 0
        else:
            obj_delta[...] = 1
    elif object_type == 'absorption_only':
        if unknown_type == 'delta_beta':
            obj_delta[...] = 0
        else:
            obj_beta[...] = 0
    if unknown_type == 'delta_beta' and non_negativity:
        obj_delta[obj_delta < 0] = 0
        obj_beta[obj_beta < 0] = 0
    elif unknown_type == 'real_imag':
        obj_delta, obj_beta = mag_phase_to_real_imag(obj_delta, obj_beta)
    return obj_delta, obj_beta


def initialize_probe(probe_size, probe_type, pupil_function=None, probe_initial=None, rescale_intensity=False,
                     prj=None, raw_data_type='magnitude', normalize_fft=False, n_probe_modes=1, sign_convention=1,
                     extra_defocus_cm=None, **kwargs):
    """adorym/util.py:198-283 for probe_type in gaussian / aperture_defocus / supplied / fixed / plane."""
    from .propagate import get_kernel
    if probe_type == 'gaussian':
        mag, phase = generate_gaussian_map(probe_size, 1, kwargs['probe_mag_sigma'], kwargs['probe_phase_max'],
                                           kwargs['probe_phase_sigma'])
        probe_real, probe_imag = mag_phase_to_real_imag(mag, phase)
    elif probe_type == 'aperture_defocus':
        mag = generate_disk(probe_size, kwargs['aperture_radius'])
        if kwargs.get('beamstop_radius', 0) > 0:
            mag = mag * (1 - generate_disk(probe_size, kwargs['beamstop_radius']))
        voxel_nm = [kwargs['psize_cm'] * 1e7] * 3
        h = get_kernel(kwargs['probe_defocus_cm'] * 1e7, kwargs['lmbda_nm'], voxel_nm, probe_size,
                       sign_convention=sign_convention)
        w = np.fft.ifft2(np.fft.fft2(mag + 0j) * h)
        probe_real, probe_imag = w.real, w.imag
    elif probe_type in ['supplied', 'fixed']:
        probe_mag, probe_phase = probe_initial
        probe_real, probe_imag = mag_phase_to_real_imag(np.asarray(probe_mag), np.asarray(probe_phase))
    elif probe_type == 'plane':
        probe_real = np.ones(probe_size)
        probe_imag = np.zeros(probe_size)
    else:
        raise ValueError('Invalid wavefront type. Choose from \'plane\', \'fixed\', \'supplied\'.')
    if pupil_function is not None:
        probe_real = probe_real * pupil_function
        probe_imag = probe_imag * pupil_function
    if extra_defocus_cm is not None:
        voxel_nm = [kwargs['psize_cm'] * 1e7] * 3
        h = get_kernel(extra_defocus_cm * 1e7, kwargs['lmbda_nm'], voxel_nm, probe_size, sign_convention=sign_convention)
        w = np.fft.ifft2(np.fft.fft2(probe_real + 1j * probe_imag) * h)
        probe_real, probe_imag = w.real, w.imag
    if rescale_intensity:
        dat = np.abs(np.asarray(prj[0:1]))
        if raw_data_type == 'magnitude':
            dat = dat ** 2
        npx = np.prod(probe_real.shape[-2:])
        if not normalize_fft:
            if sign_convention == 1:
                target = np.sum(np.mean(np.abs(dat), axis=(0, 1))) / npx
            else:
                target = np.sum(np.mean(np.abs(dat), axis=(0, 1))) * npx
        else:
            target = np.sum(np.mean(np.abs(dat), axis=(0, 1)))
        current = np.sum(probe_real ** 2 + probe_imag ** 2)
        s = np.sqrt(target / current)
        probe_real, probe_imag = probe_real * s, probe_imag * s
    return probe_real, probe_imag


def split_tasks(arr, split_size):
    """adorym/util.py:1629-1635."""
    return [arr[i:min(i + split_size, len(arr))] for i in range(0, len(arr), split_size)]
