"""Array containers of the data-parallel mode (adorym/array_ops.py:165-345), device resident."""
import numpy as np
import torch

from .util import initialize_object_for_dp


class LargeArray(object):
    def __init__(self, full_size, distribution_mode=None, monochannel=False, output_folder=None, device=None):
        if distribution_mode is not None:
            raise NotImplementedError('adorym_b200: distribution_mode=%r is out of scope' % (distribution_mode,))
        self.full_size = full_size
        self.distribution_mode = distribution_mode
        self.monochannel = monochannel
        self.output_folder = output_folder
        self.arr = None
        self.arr_rot = None
        self.device = device


class ObjectFunction(LargeArray):
    def __init__(self, full_size, distribution_mode=None, output_folder=None, ds_level=1, object_type='normal', device=None):
        super(ObjectFunction, self).__init__(full_size, distribution_mode=distribution_mode, monochannel=False,
                                             output_folder=output_folder, device=device)
        self.ds_level = ds_level
        self.object_type = object_type

    def initialize_array(self, save_stdout=None, timestr=None, not_first_level=False, initial_guess=None, device=None,
                         random_guess_means_sigmas=(8.7e-7, 5.1e-8, 1e-7, 1e-8), unknown_type='delta_beta',
                         non_negativity=False, seed=None):
        d, b = initialize_object_for_dp(self.full_size[:-1], object_type=self.object_type, initial_guess=initial_guess,
                                        random_guess_means_sigmas=random_guess_means_sigmas, unknown_type=unknown_type,
                                        non_negativity=non_negativity, seed=seed)
        self.initialize_array_with_values(d, b, device=device)

    def initialize_array_with_values(self, obj_delta, obj_beta, device=None):
        a = np.stack([np.asarray(obj_delta), np.asarray(obj_beta)], -1).astype('float32')
        self.arr = torch.from_numpy(a).to(device).requires_grad_(True)

    def apply_finite_support_mask_to_array(self, mask, unknown_type='delta_beta', device=None):
        """adorym/array_ops.py:239-251."""
        with torch.no_grad():
            m = mask.mask.unsqueeze(-1)
            if unknown_type == 'delta_beta':
                new = self.arr * m
            else:
                vac = torch.tensor([1., 0.], device=self.arr.device)
                new = self.arr * m + vac * (1 - m)
        self.arr = new.requires_grad_(True)


class Gradient(ObjectFunction):
    def __init__(self, obj, forward_model=None):
        super(Gradient, self).__init__(obj.full_size, obj.distribution_mode, obj.output_folder, obj.ds_level, obj.object_type)
        self.forward_model = forward_model

    def initialize_array_with_values(self, obj_delta, obj_beta, device=None):
        a = np.stack([np.asarray(obj_delta), np.asarray(obj_beta)], -1).astype('float32')
        self.arr = torch.from_numpy(a).to(device)


class Mask(LargeArray):
    def __init__(self, full_size, finite_support_mask_path, distribution_mode=None, output_folder=None, ds_level=1):
        super(Mask, self).__init__(full_size, distribution_mode, monochannel=True, output_folder=output_folder)
        self.mask = None
        self.ds_level = ds_level
        self.finite_support_mask_path = finite_support_mask_path

    def initialize_array_with_values(self, mask, device=None, dtype=None):
        self.mask = torch.as_tensor(np.asarray(mask), dtype=torch.float32).to(device).contiguous()

    def update_mask_array(self, obj, threshold=1e-9):
        if obj.arr is not None:
            self.mask[obj.arr[..., 0] < threshold] = 0
