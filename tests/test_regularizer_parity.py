"""The regularisers ALONE (data term off): value and gradient of L1 / TV / L1 + TV / reweighted L1 (+ TV) on a small
non-cubic object with negative values and exact zeros, against the unmodified reference's golden
(tests/golden/make_golden_reg.py -> reg.npz).  In the end-to-end cases of the reference's demo configurations the
regulariser contributes 1e-8 .. 1e-6 of the gradient, so those cannot see an error in it; here it is the whole gradient.
Tolerance 1e-5 (fp32 sums of a few thousand terms)."""
import ctypes
import os
import sys
import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
from make_golden_reg import COMBOS, reg_inputs  # noqa: E402
from oracle import multislice_oracle as orc  # noqa: E402
from oracle_util import rel_err  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reg.npz'))
TOL = 1e-5


@pytest.mark.parametrize('name', sorted(COMBOS))
def test_oracle_regularizers_match_reference_golden(name):
    c = COMBOS[name]
    obj, w = reg_inputs()
    x = torch.tensor(obj, requires_grad=True)
    val = orc.regularization_value(x, alpha_d=c['alpha_d'], alpha_b=c['alpha_b'], gamma=c['gamma'],
                                   weight_l1=torch.tensor(w) if c['rw'] else None)
    (g,) = torch.autograd.grad(val, [x])
    assert abs(val.item() - float(GOLD[name + '_value'])) <= TOL * abs(float(GOLD[name + '_value']))
    assert rel_err(g.numpy(), GOLD[name + '_grad']) < TOL


@pytest.mark.parametrize('name', sorted(COMBOS))
@pytest.mark.parametrize('slabs', [1, 3])
def test_emulated_regularizer_kernel_matches_reference_golden(name, slabs):
    """k_reg_l1tv compiled for the host (tests/emu): value, gradient written (accumulate = 0, scaled by a device
    scalar) on y slabs, and gradient accumulated (accumulate = 1)."""
    import emu_lib
    L = emu_lib.emu()
    P = emu_lib.P
    c = COMBOS[name]
    obj, w = reg_inputs()
    Y, X, Z, _ = obj.shape
    wl = np.ascontiguousarray(w) if c['rw'] else None
    ad, ab, gm = float(c['alpha_d'] or 0.), float(c['alpha_b'] or 0.), float(c['gamma'] or 0.)
    val = np.zeros((1,), 'float32')
    assert L.adm_regularizer_l1tv_scaled(Y, X, Z, P(obj), P(wl), ad, ab, gm, P(val), None, None, 0, None) == 0
    assert abs(val[0] - float(GOLD[name + '_value'])) <= TOL * abs(float(GOLD[name + '_value']))
    g = np.full(obj.shape, np.nan, 'float32')
    gs = np.array([2.5], 'float32')
    edges = [int(round(i * Y / slabs)) for i in range(slabs + 1)]
    for y0, y1 in zip(edges[:-1], edges[1:]):
        assert L.adm_regularizer_l1tv_range(Y, X, Z, P(obj), P(wl), ad, ab, gm, None, P(g), P(gs), 0, y0, y1, None) == 0
    assert rel_err(g, 2.5 * GOLD[name + '_grad']) < TOL
    base = np.random.RandomState(0).normal(size=obj.shape).astype('float32') * 1e-3
    g2 = base.copy()
    assert L.adm_regularizer_l1tv(Y, X, Z, P(obj), P(wl), ad, ab, gm, None, P(g2), None) == 0
    assert rel_err(g2 - base, GOLD[name + '_grad']) < 1e-4      # (rounding of base + g in fp32)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(COMBOS))
def test_gpu_regularizers_match_reference_golden(name):
    """The public regulariser classes (adorym/regularizers.py signatures) and the fused RegL1TVFn on the B200."""
    import adorym_b200 as ab
    from adorym_b200 import ops
    dev = torch.device('cuda', 0)
    c = COMBOS[name]
    obj, w = reg_inputs()
    x = torch.tensor(obj, device=dev, requires_grad=True)
    regs = []
    if c['alpha_d'] is not None:
        if c['rw']:
            r = ab.ReweightedL1Regularizer(c['alpha_d'], c['alpha_b'])
            r.update_l1_weight(torch.tensor(w, device=dev))
        else:
            r = ab.L1Regularizer(c['alpha_d'], c['alpha_b'])
        regs.append(r)
    if c['gamma'] is not None:
        regs.append(ab.TVRegularizer(c['gamma']))
    val = sum(r.get_value(x, device=dev) for r in regs)
    (g,) = torch.autograd.grad(val, [x])
    assert abs(val.item() - float(GOLD[name + '_value'])) <= TOL * abs(float(GOLD[name + '_value']))
    assert rel_err(g.cpu().numpy(), GOLD[name + '_grad']) < TOL
    # fused: one kernel pass for all terms, upstream gradient applied on the device
    x2 = torch.tensor(obj, device=dev, requires_grad=True)
    v2 = ops.RegL1TVFn.apply(x2, torch.tensor(w, device=dev) if c['rw'] else None, c['alpha_d'], c['alpha_b'], c['gamma'])
    (g2,) = torch.autograd.grad(3.0 * v2, [x2])
    assert abs(v2.item() - float(GOLD[name + '_value'])) <= TOL * abs(float(GOLD[name + '_value']))
    assert rel_err(g2.cpu().numpy(), 3.0 * GOLD[name + '_grad']) < TOL
