"""Optimizers with the reference's constructor / apply_gradient signatures (adorym/optimizers.py).
Adam runs as ONE fused kernel (adm_adam_step) instead of ~12 elementwise launches."""
import numpy as np
import torch

from . import ops
from .array_ops import Gradient


class Optimizer(object):
    """adorym/optimizers.py:32-252 (data-parallel mode only)."""

    def __init__(self, name, output_folder='.', params_list=(), distribution_mode=None, options_dict=None, forward_model=None):
        if distribution_mode is not None:
            raise NotImplementedError('adorym_b200: distribution_mode=%r is out of scope' % (distribution_mode,))
        self.name = name
        self.forward_model = forward_model
        self.output_folder = output_folder
        self.params_list = params_list
        self.params_whole_array_dict = {}
        self.i_batch = 0
        self.index_in_grad_returns = None
        self.distribution_mode = distribution_mode
        self.options_dict = options_dict if options_dict is not None else {}
        self.grads = None
        self.whole_object_size = None

    def __str__(self):
        s = self.__class__.__name__ + '; '
        for k in self.options_dict.keys():
            s = s + k + ': ' + str(self.options_dict[k]) + '; '
        return s

    def create_container(self, whole_object_size, use_checkpoint, device_obj, use_numpy=False, dtype='float32'):
        self.create_param_arrays(whole_object_size, device=device_obj)

    def create_param_arrays(self, whole_object_size, device=None, use_numpy=False):
        self.whole_object_size = list(whole_object_size)
        for p in self.params_list:
            self.params_whole_array_dict[p] = torch.zeros(self.whole_object_size, dtype=torch.float32, device=device)

    def set_index_in_grad_return(self, ind):
        self.index_in_grad_returns = ind

    def convert_gradient(self, gradient):
        return gradient.arr if isinstance(gradient, Gradient) else gradient


class AdamOptimizer(Optimizer):
    """adorym/optimizers.py:268-337."""

    def __init__(self, name, output_folder='.', distribution_mode=None, options_dict=None, forward_model=None):
        super(AdamOptimizer, self).__init__(name, output_folder=output_folder, params_list=['m', 'v'],
                                            distribution_mode=distribution_mode, options_dict=options_dict,
                                            forward_model=forward_model)

    def apply_gradient(self, x, gradient, i_batch, step_size=0.001, b1=0.9, b2=0.999, eps=1e-7, distribution_mode=False,
                       m=None, v=None, return_moments=False, update_batch_count=True, use_numpy=False,
                       params_slicer=None, **kwargs):
        """Same update as optimizers.py:309-318.  `i_batch` is the user-specified step number (the driver passes
        i_opt_batch, which by default advances per angle, quirk 4).  The update is done in place on the storage of
        `x`, `m`, `v` by one fused kernel; the returned tensor is `x` detached (the reference returns a new tensor
        which the driver immediately re-attaches, ptychography.py:1125-1129).  Optional fused epilogue kwargs:
        clip_nonneg (bool), zero_channel (0 | 1 | None), mask (per-voxel multiplier tensor)."""
        if use_numpy or distribution_mode or params_slicer is not None:
            raise NotImplementedError('adorym_b200.AdamOptimizer: numpy / distributed / sliced updates are out of scope')
        g = self.convert_gradient(gradient)
        if m is None or v is None:
            if 'm' not in self.params_whole_array_dict or self.params_whole_array_dict['m'].shape != x.shape:
                self.create_param_arrays(list(x.shape), device=x.device)
            m = self.params_whole_array_dict['m']
            v = self.params_whole_array_dict['v']
        with torch.no_grad():
            xd = x.detach()
            if not xd.is_contiguous():
                raise RuntimeError('adorym_b200.AdamOptimizer: x must be contiguous')
            ops.adam_step_(xd, m, v, g.detach(), int(i_batch), step_size=step_size, b1=b1, b2=b2, eps=eps,
                           clip_nonneg=kwargs.get('clip_nonneg', False), zero_channel=kwargs.get('zero_channel', None),
                           mask=kwargs.get('mask', None))
        if update_batch_count:
            self.i_batch += 1
        if return_moments:
            return xd, m, v
        return xd


class GDOptimizer(Optimizer):
    """adorym/optimizers.py:432-464; the update runs as one fused kernel (adm_sgd_step) in place on x."""

    def __init__(self, name, output_folder='.', distribution_mode=None, options_dict=None, forward_model=None):
        super(GDOptimizer, self).__init__(name, output_folder=output_folder, params_list=[],
                                          distribution_mode=distribution_mode, options_dict=options_dict,
                                          forward_model=forward_model)

    def apply_gradient(self, x, gradient, i_batch, step_size=0.001, dynamic_rate=True, first_downrate_iteration=92,
                       use_numpy=False, verbose=False, **kwargs):
        if use_numpy:
            raise NotImplementedError('adorym_b200.GDOptimizer: numpy updates are out of scope')
        g = self.convert_gradient(gradient)
        if dynamic_rate:
            threshold_iteration = first_downrate_iteration
            i = 1
            while threshold_iteration < i_batch:
                threshold_iteration += first_downrate_iteration * 2 ** i
                i += 1
                step_size /= 2.
        with torch.no_grad():
            xd = x.detach()
            ops.sgd_step_(xd, None, g.detach(), step_size=step_size, clip_nonneg=kwargs.get('clip_nonneg', False),
                          zero_channel=kwargs.get('zero_channel', None), mask=kwargs.get('mask', None))
        return xd


class MomentumOptimizer(Optimizer):
    """adorym/optimizers.py:366-411: v <- gamma v + step g, x <- x - v (fused kernel, in place on x and v)."""

    def __init__(self, name, output_folder='.', distribution_mode=None, options_dict=None, forward_model=None):
        super(MomentumOptimizer, self).__init__(name, output_folder=output_folder, params_list=['v'],
                                                distribution_mode=distribution_mode, options_dict=options_dict,
                                                forward_model=forward_model)

    def apply_gradient(self, x, gradient, i_batch, step_size=0.001, gamma=0.9, use_numpy=False, params_slicer=None,
                       **kwargs):
        if use_numpy or params_slicer is not None:
            raise NotImplementedError('adorym_b200.MomentumOptimizer: numpy / sliced updates are out of scope')
        g = self.convert_gradient(gradient)
        if 'v' not in self.params_whole_array_dict or self.params_whole_array_dict['v'].shape != x.shape:
            self.create_param_arrays(list(x.shape), device=x.device)
        v = self.params_whole_array_dict['v']
        with torch.no_grad():
            xd = x.detach()
            ops.sgd_step_(xd, v, g.detach(), step_size=step_size, gamma=gamma, clip_nonneg=kwargs.get('clip_nonneg', False),
                          zero_channel=kwargs.get('zero_channel', None), mask=kwargs.get('mask', None))
        return xd


# ------------------------------------------------------------------------------------------------------
# parameter plumbing                                                     adorym/optimizers.py:793-1089
# ------------------------------------------------------------------------------------------------------
# (switch, optimizer kwarg, learning-rate kwarg, parameter name) in the order of adorym/optimizers.py:793-958: the
# gradient tuple of the Differentiator follows this order (quirk 17)
_PARAM_TABLE = (('optimize_probe_pos_offset', 'optimizer_probe_pos_offset', 'probe_pos_offset_learning_rate', 'probe_pos_offset'),
                ('optimize_prj_pos_offset', 'optimizer_prj_pos_offset', 'prj_pos_offset_learning_rate', 'prj_pos_offset'),
                ('optimize_all_probe_pos', 'optimizer_all_probe_pos', 'all_probe_pos_learning_rate', 'probe_pos_correction'),
                ('optimize_slice_pos', 'optimizer_slice_pos', 'slice_pos_learning_rate', 'slice_pos_cm_ls'),
                ('optimize_free_prop', 'optimizer_free_prop', 'free_prop_learning_rate', 'free_prop_cm'),
                ('optimize_tilt', 'optimizer_tilt', 'tilt_learning_rate', 'tilt_ls'),
                ('optimize_prj_affine', 'optimizer_prj_affine', 'prj_affine_learning_rate', 'prj_affine_ls'))
# optimize_probe_defocusing: the reference's PyTorch backend itself fails on it (numpy get_kernel on a tensor that
# requires grad, forward_model.py:247-251); the others have no accelerated forward path
_UNSUPPORTED_PARAM_SWITCHES = ('optimize_probe_defocusing',
                               'optimize_ctf_lg_kappa')


def create_and_initialize_parameter_optimizers(optimizable_params, kwargs):
    """adorym/optimizers.py:793-958 for the refinable parameters on the accelerated path: probe, per-angle probe position
    offset, probe positions (MultiDist: per-distance data shift), propagation distances, tilt angles, affine alignment of
    the data.  The gradient tuple is ordered (obj, probe_real, probe_imag, <the rest in the order above>) (quirk 17)."""
    for sw in _UNSUPPORTED_PARAM_SWITCHES:
        if kwargs.get(sw, False):
            raise NotImplementedError('adorym_b200: %s is not on the accelerated path' % sw)
    opt_ls = kwargs['opt_ls']
    forward_model = kwargs['forward_model']
    output_folder = kwargs['output_folder']
    device_obj = kwargs['device_obj']
    opt_args_ls = [0]
    if kwargs['optimize_probe']:
        if kwargs['optimizer_probe'] is not None:
            opt_probe = kwargs['optimizer_probe']
            opt_probe.name = 'probe'
        else:
            opt_probe = AdamOptimizer('probe', output_folder=output_folder,
                                      options_dict={'step_size': kwargs['probe_learning_rate']}, forward_model=forward_model)
        opt_probe.create_param_arrays([kwargs['n_probe_modes'], *kwargs['probe_size'], 2], device=device_obj)
        opt_probe.set_index_in_grad_return(len(opt_args_ls))
        opt_args_ls = opt_args_ls + [forward_model.get_argument_index('probe_real'),
                                     forward_model.get_argument_index('probe_imag')]
        opt_ls.append(opt_probe)
    for switch, opt_kw, lr_kw, name in _PARAM_TABLE:
        if not kwargs.get(switch, False):
            continue
        if name == 'free_prop_cm' and not kwargs.get('is_multi_dist', False):      # optimizers.py:905
            continue
        if name == 'slice_pos_cm_ls' and not kwargs.get('is_sparse_multislice', False):      # optimizers.py:891
            continue
        if name == 'probe_pos_offset':
            assert kwargs['optimize_all_probe_pos'] == False                       # optimizers.py:846
        if kwargs.get(opt_kw) is not None:
            o = kwargs[opt_kw]
            o.name = name
        elif name in ('probe_pos_offset', 'prj_pos_offset'):
            # optimizers.py:851-857, :868-871: plain gradient descent without the step-halving schedule
            o = GDOptimizer(name, output_folder=output_folder,
                            options_dict={'step_size': kwargs[lr_kw], 'dynamic_rate': False})
        else:
            o = AdamOptimizer(name, output_folder=output_folder, options_dict={'step_size': kwargs[lr_kw]},
                              forward_model=forward_model)
        o.create_param_arrays(list(optimizable_params[name].shape), device=device_obj)
        o.set_index_in_grad_return(len(opt_args_ls))
        opt_args_ls.append(forward_model.get_argument_index(name))
        opt_ls.append(o)
    return opt_ls, opt_args_ls


def initialize_parameter_gradients(opt_ls, device=None, use_numpy=False):
    for opt in opt_ls:
        if opt.name != 'obj':
            opt.grads = torch.zeros(opt.whole_object_size, dtype=torch.float32, device=device)
    return opt_ls


def update_parameter_gradients(opt_ls, grads, use_numpy=False):
    """adorym/optimizers.py:974-990 (probe gradient = stack(grads[1:3]), quirk 17)."""
    for opt in opt_ls:
        if opt.name == 'obj':
            continue
        elif opt.name == 'probe':
            opt.grads += torch.stack(list(grads[1:3]), dim=-1)
        else:
            opt.grads += grads[opt.index_in_grad_returns]
    return opt_ls


def update_parameters(opt_ls, optimizable_params, kwargs):
    """adorym/optimizers.py:993-1089: SUM all-reduce of each parameter gradient, Adam, drift removal (quirk 16)."""
    i_epoch, i_batch, n_batch = kwargs['i_epoch'], kwargs['i_batch'], kwargs['n_batch']
    other_params_update_delay = kwargs['other_params_update_delay']
    probe_update_delay = kwargs['probe_update_delay']
    probe_update_limit = kwargs['probe_update_limit']
    i_full_angle = kwargs['i_opt_batch']
    comm = kwargs['comm']
    if probe_update_limit is None:
        probe_update_limit = np.inf
    it = i_batch + i_epoch * n_batch
    for opt in opt_ls:
        if opt.name == 'obj':
            continue
        if opt.name == 'probe':
            if probe_update_delay <= it < probe_update_limit:
                with torch.no_grad():
                    opt.grads = comm.allreduce(opt.grads)
                    probe = torch.stack([optimizable_params['probe_real'], optimizable_params['probe_imag']], dim=-1).contiguous()
                    probe = opt.apply_gradient(probe, opt.grads, i_full_angle, **opt.options_dict)
                    optimizable_params['probe_real'] = probe[..., 0].contiguous().requires_grad_(True)
                    optimizable_params['probe_imag'] = probe[..., 1].contiguous().requires_grad_(True)
                    opt.grads = None
        elif it >= other_params_update_delay:
            if opt.name == 'probe_pos_correction':
                with torch.no_grad():
                    opt.grads = comm.allreduce(opt.grads)
                    ppc = optimizable_params['probe_pos_correction'].detach().clone().contiguous()
                    ppc = opt.apply_gradient(ppc, opt.grads, i_full_angle, **opt.options_dict)
                    dims = tuple(range(ppc.dim() - 1))
                    ppc = ppc - ppc.mean(dim=dims)
                    optimizable_params['probe_pos_correction'] = ppc.requires_grad_(True)
            elif opt.name == 'slice_pos_cm_ls':
                # adorym/optimizers.py:1050-1060: Adam, then the first slice is pinned to position 0
                with torch.no_grad():
                    opt.grads = comm.allreduce(opt.grads)
                    sp = optimizable_params['slice_pos_cm_ls'].detach().clone().contiguous()
                    sp = opt.apply_gradient(sp, opt.grads, i_full_angle, **opt.options_dict)
                    optimizable_params['slice_pos_cm_ls'] = (sp - sp[0]).requires_grad_(True)
            elif opt.name == 'prj_affine_ls':
                # adorym/optimizers.py:1062-1075: Adam, then the transformation of the first frame is pinned to identity
                with torch.no_grad():
                    opt.grads = comm.allreduce(opt.grads)
                    aff = optimizable_params['prj_affine_ls'].detach().clone().contiguous()
                    aff = opt.apply_gradient(aff, opt.grads, i_full_angle, **opt.options_dict)
                    aff[0] = torch.tensor([[1., 0., 0.], [0., 1., 0.]], dtype=aff.dtype, device=aff.device)
                    optimizable_params['prj_affine_ls'] = aff.requires_grad_(True)
            else:
                # adorym/optimizers.py:1077-1083: plain Adam on any other refinable parameter (free_prop_cm, ...)
                with torch.no_grad():
                    opt.grads = comm.allreduce(opt.grads)
                    var = optimizable_params[opt.name].detach().clone().contiguous()
                    var = opt.apply_gradient(var, opt.grads, i_full_angle, **opt.options_dict)
                    optimizable_params[opt.name] = var.requires_grad_(True)
    return optimizable_params
