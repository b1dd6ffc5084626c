"""Regularizer classes with the reference's constructors (adorym/regularizers.py).  For delta_beta objects the
value and its gradient come from one fused kernel (adm_regularizer_l1tv)."""
import torch

from . import ops


def mismatch_loss_from_magnitude(pred, prj_abs, loss_function_type='lsq', raw_data_type='magnitude',
                                 poisson_multiplier=1., mask=None):
    """adorym/forward_model.py:88-103 (+ beamstop :128-136) on magnitudes -- API-compatibility path (library ops)."""
    if mask is not None:
        sel = mask.bool().unsqueeze(0).expand_as(pred)
        pred = pred[sel].reshape(pred.shape[0], -1)
        prj_abs = prj_abs[sel].reshape(prj_abs.shape[0], -1)
    pm = poisson_multiplier
    if loss_function_type == 'lsq':
        m = prj_abs if raw_data_type == 'magnitude' else torch.sqrt(prj_abs)
        return torch.mean((pred - m) ** 2)
    i_meas = prj_abs ** 2 if raw_data_type == 'magnitude' else prj_abs
    return torch.mean(pred ** 2 * pm - i_meas * pm * torch.log(pred ** 2 * pm))


class Regularizer(object):
    def __init__(self, unknown_type='delta_beta'):
        self.unknown_type = unknown_type

    def get_value(self, obj, device=None, **kwargs):
        pass


def _tv3d(a):
    r = (torch.roll(a, 1, 0) - a).abs().sum() + (torch.roll(a, 1, 1) - a).abs().sum() + (torch.roll(a, 1, 2) - a).abs().sum()
    return r / a.numel()


class L1Regularizer(Regularizer):
    """adorym/regularizers.py:18-46."""

    def __init__(self, alpha_d, alpha_b, unknown_type='delta_beta'):
        super().__init__(unknown_type)
        self.alpha_d = alpha_d
        self.alpha_b = alpha_b

    def get_value(self, obj, device=None, **kwargs):
        if self.unknown_type == 'delta_beta':
            return ops.RegL1TVFn.apply(obj, None, self.alpha_d, self.alpha_b, 0.)
        r, i = obj[..., 0], obj[..., 1]
        reg = torch.zeros((), device=obj.device)
        if self.alpha_d not in [None, 0]:
            om = torch.sqrt(r ** 2 + i ** 2)
            reg = reg + self.alpha_d * torch.mean(torch.abs(om - torch.mean(om)))
        if self.alpha_b not in [None, 0]:
            reg = reg + self.alpha_b * torch.mean(torch.abs(torch.atan2(i, r)))
        return reg


class ReweightedL1Regularizer(Regularizer):
    """adorym/regularizers.py:49-83."""

    def __init__(self, alpha_d, alpha_b, unknown_type='delta_beta'):
        super().__init__(unknown_type)
        self.alpha_d = alpha_d
        self.alpha_b = alpha_b
        self.weight_l1 = None

    def update_l1_weight(self, weight_l1):
        """regularizers.py:60-61.  New weights of the same shape are copied INTO the existing tensor, so that launches
        captured in a CUDA graph (adorym_b200.graph) keep reading the current weights."""
        if weight_l1 is getattr(self, '_weight_src', None):
            return
        self._weight_src = weight_l1
        w = self.weight_l1
        if (torch.is_tensor(w) and torch.is_tensor(weight_l1) and w.shape == weight_l1.shape and w.dtype == weight_l1.dtype
                and w.device == weight_l1.device and w.data_ptr() != weight_l1.data_ptr()):
            with torch.no_grad():
                w.copy_(weight_l1)
        else:
            self.weight_l1 = weight_l1.clone() if torch.is_tensor(weight_l1) else weight_l1

    def get_value(self, obj, device=None, **kwargs):
        if self.unknown_type == 'delta_beta':
            return ops.RegL1TVFn.apply(obj, self.weight_l1, self.alpha_d, self.alpha_b, 0.)
        r, i = obj[..., 0], obj[..., 1]
        wm = self.weight_l1[..., 0] ** 2 + self.weight_l1[..., 1] ** 2
        reg = torch.zeros((), device=obj.device)
        if self.alpha_d not in [None, 0]:
            om = torch.sqrt(r ** 2 + i ** 2)
            reg = reg + self.alpha_d * torch.mean(wm * torch.abs(om - torch.mean(om)))
        if self.alpha_b not in [None, 0]:
            reg = reg + self.alpha_b * torch.mean(wm * torch.abs(torch.atan2(i, r)))
        return reg


class TVRegularizer(Regularizer):
    """adorym/regularizers.py:86-110."""

    def __init__(self, gamma, unknown_type='delta_beta'):
        super().__init__(unknown_type)
        self.gamma = gamma

    def get_value(self, obj, distribution_mode=None, device=None, **kwargs):
        if self.unknown_type == 'delta_beta':
            return ops.RegL1TVFn.apply(obj, None, 0., 0., self.gamma)
        r, i = obj[..., 0], obj[..., 1]
        return self.gamma * _tv3d(r ** 2 + i ** 2) + self.gamma * _tv3d(torch.atan2(i, r))
