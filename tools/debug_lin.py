import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import numpy as np, torch
import adorym_b200 as ab
from adorym_b200 import ops
from oracle import multislice_oracle as orc
from oracle_util import rel_err
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(0)
Y = X = Z = 256
P = 72
obj = torch.zeros((Y, X, Z, 2), device=dev)
obj[..., 0] = torch.rand((Y, X, Z), device=dev, generator=g) * 1e-5
pos = np.array([(y, x) for y in (np.arange(23) * 12) - 36 for x in (np.arange(23) * 12) - 36], dtype='int32')
pos_t = torch.tensor(pos, device=dev)
voxel_nm = np.array([1e-7] * 3) * 1e7
h = ab.get_kernel(voxel_nm[-1], 1240. / 5000., voxel_nm, [P, P])
plan = ops.MultislicePlan(P, Z, Y, X, ops.k1_of(5000., 1e-7), h, dev, free_prop='inf')
lut = orc.rotation_coords([Y, X, Z], np.float32(0.4))
rot = ops.RotateFn.apply(obj, ops.normalized_grid(lut, X, Z, dev))
pr, pi = orc.gaussian_probe([P, P], 1, 6, 0.5, 6)
probe = torch.tensor(np.stack([pr, pi], -1)[None].astype('float32'), device=dev)
y = torch.randn((529, 1, P, P, 2), device=dev, generator=g)
rot2 = rot.detach().requires_grad_(True)
def grad(sub, yy):
    d = ops.MultisliceFn.apply(rot2, probe, sub, plan)
    (g1,) = torch.autograd.grad((d * yy).sum(), [rot2])
    return g1.double()
for lo, hi in ((100, 140), (0, 1), (100, 101), (264, 265)):
    sub = pos_t[lo:hi].contiguous()
    ga = grad(sub, y[lo:hi]); gb = grad(sub, y[lo:hi])
    g2 = grad(torch.cat([sub, sub]).contiguous(), torch.cat([y[lo:hi], y[lo:hi]]))
    n = ga.norm().item()
    print(lo, hi, 'repeat diff', ((ga - gb).norm() / n).item(), 'lin diff', ((g2 - 2 * ga).norm() / n).item(),
          'max|g|', ga.abs().max().item(), 'nan', torch.isnan(ga).any().item())
    d = (g2 - 2 * ga).abs()
    idx = torch.nonzero(d > 1e-3 * ga.abs().max())
    print('  n bad', idx.shape[0], idx[:5].tolist())
print('--- adjoint test wrt probe')
for lo, hi in ((0, 1), (100, 140), (0, 529)):
    sub = pos_t[lo:hi].contiguous()
    pp = probe.clone().requires_grad_(True)
    det = ops.MultisliceFn.apply(rot, pp, sub, plan)
    yy = y[lo:hi]
    (gp,) = torch.autograd.grad((det * yy).sum(), [pp])
    lhs = (det.double() * yy.double()).sum().item()
    rhs = (pp.double() * gp.double()).sum().item()
    # per-pattern probe path (no atomics)
    ppb = probe[None].repeat(hi - lo, 1, 1, 1, 1).clone().requires_grad_(True)
    det2 = ops.MultisliceFn.apply(rot, ppb, sub, plan)
    (gpb,) = torch.autograd.grad((det2 * yy).sum(), [ppb])
    rhs2 = (ppb.double() * gpb.double()).sum().item()
    print(lo, hi, 'lhs', lhs, 'rhs', rhs, 'rel', abs(lhs - rhs) / abs(lhs), 'rhs(per-pattern)', rhs2,
          'det diff', (det - det2).abs().max().item(), 'gp vs sum gpb', ((gp.double() - gpb.double().sum(0)).norm() / gp.double().norm()).item())
