"""Golden vectors of the regularisers alone (data term off), from the UNMODIFIED reference (adorym/regularizers.py,
PyTorch backend on CPU): value and full gradient of L1, TV, L1 + TV, reweighted L1 and reweighted L1 + TV on a small
non-cubic delta/beta object that contains negative values and exact zeros.  Run in the build container only:

    PYTHONPATH=oracle/ref_shims:/root/reference python tests/golden/make_golden_reg.py
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

COMBOS = {
    'l1': dict(alpha_d=0.7, alpha_b=0.3, gamma=None, rw=False),
    'tv': dict(alpha_d=None, alpha_b=None, gamma=0.9, rw=False),
    'l1_tv': dict(alpha_d=0.7, alpha_b=0.3, gamma=0.9, rw=False),
    'rwl1': dict(alpha_d=0.5, alpha_b=0.2, gamma=None, rw=True),
    'rwl1_tv': dict(alpha_d=0.5, alpha_b=0.2, gamma=1.3, rw=True),
}


def reg_inputs():
    rs = np.random.RandomState(41)
    shape = (10, 12, 14, 2)
    obj = rs.normal(size=shape).astype('float32') * np.array([1e-3, 1e-4], 'float32')
    obj[rs.uniform(size=shape) < 0.05] = 0.0            # exact zeros: sign(0) = 0 subgradient
    obj[2, 3, 4:9, :] = obj[2, 3, 4, :]                  # equal neighbours: TV subgradient at 0
    w = rs.uniform(0.2, 3.0, size=shape).astype('float32')
    return obj, w


def main():
    import torch
    import adorym  # noqa: F401
    import adorym.global_settings as gs
    from adorym.regularizers import L1Regularizer, ReweightedL1Regularizer, TVRegularizer
    gs.backend = 'pytorch'
    obj, w = reg_inputs()
    out = {}
    for name, c in COMBOS.items():
        x = torch.tensor(obj, requires_grad=True)
        regs = []
        if c['alpha_d'] is not None:
            if c['rw']:
                r = ReweightedL1Regularizer(c['alpha_d'], c['alpha_b'])
                r.update_l1_weight(torch.tensor(w))
            else:
                r = L1Regularizer(c['alpha_d'], c['alpha_b'])
            regs.append(r)
        if c['gamma'] is not None:
            regs.append(TVRegularizer(c['gamma']))
        val = sum(r.get_value(x, device=None) for r in regs)
        (g,) = torch.autograd.grad(val, [x])
        out[name + '_value'] = np.float64(val.item())
        out[name + '_grad'] = g.numpy()
        print(name, float(val))
    np.savez_compressed(os.path.join(HERE, 'reg.npz'), **out)


if __name__ == '__main__':
    main()
