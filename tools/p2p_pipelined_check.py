"""debug: pipelined vs plain peer update on synthetic gradients (torchrun, N ranks)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adorym_b200 as ab
from adorym_b200 import dist as adist, ops
rank, world, local = adist.init_from_env('nccl')
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
shape = [64, 48, 40, 2]
peer = adist.PeerUpdate(shape, dev)
torch.manual_seed(0)
x0 = torch.randn(shape, device=dev) * 1e-3
torch.manual_seed(1 + rank)
g = torch.randn(shape, device=dev)
def run(pipelined):
    with torch.no_grad():
        peer.object_tensor().copy_(x0)
        peer.gradient_tensor().copy_(g)
    torch.cuda.synchronize(); torch.distributed.barrier()
    opt = ab.AdamOptimizer('obj', options_dict={'step_size': 1e-4})
    if pipelined:
        slabs = []
        for (y0, y1) in ops.slab_ranges(shape[0], 4):
            ev = torch.cuda.Event(); ev.record(); slabs.append((y0, y1, ev))
        x = peer.update_pipelined(opt, 0, slabs, step_size=1e-4)
    else:
        x = peer.update(opt, 0, step_size=1e-4)
    torch.cuda.synchronize(); torch.distributed.barrier()
    return x.detach().clone(), opt.params_whole_array_dict['m'].clone()
xa, ma = run(False)
xb, mb = run(True)
print(rank, 'nan a', int(torch.isnan(xa).sum()), 'nan b', int(torch.isnan(xb).sum()), 'max diff x', float((xa - xb).abs().max()),
      'max diff m (own shard differs by layout)', float((ma - mb).abs().max()), flush=True)
