#!/usr/bin/env python
"""Multi-GPU check of reconstruct_ptychography (run under torchrun on N GPUs of one box):
the fused NVLink peer-memory update (default) and the NCCL all-reduce path (ADM_DP=nccl) must give the same loss history
and final object (they differ only in the fp32 summation order of the gradient all-reduce), and every rank must hold the
same object.  Uses the driver-golden inputs (tests/golden/driver_cases.py).  Prints one JSON line on rank 0."""
import json
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import adorym_b200 as ab  # noqa: E402
from adorym_b200 import dist as adist  # noqa: E402
from driver_cases import driver_params, driver_mask  # noqa: E402

rank, world, local = adist.init_from_env('nccl')
torch.cuda.set_device(local)
data = np.load(os.path.join(ROOT, 'tests', 'golden', 'driver_data.npz'))['data'].astype('float32')
out = {}
res = {}
for mode in ('p2p', 'nccl', 'nccl1'):
    os.environ['ADM_DP'] = mode[:4]
    os.environ['ADM_GRAD_SLABS'] = '1' if mode == 'nccl1' else '4'
    params = driver_params('driver_immediate')
    r = ab.reconstruct_ptychography(data, **params, finite_support_mask_path=driver_mask(), verbose=False, save_outputs=False,
                                    output_folder='/tmp/adm_dist_check_%s_%d' % (mode, rank))
    res[mode] = r
    ref = r['obj'].clone()
    torch.distributed.broadcast(ref, src=0)
    d = (ref - r['obj']).abs().max().reshape(1)
    torch.distributed.all_reduce(d, op=torch.distributed.ReduceOp.MAX)
    out[mode + '_replica_max_abs_diff'] = float(d.item())
    out[mode + '_losses'] = [h[2] for h in r['loss_history']]
lp, ln = np.array(out['p2p_losses']), np.array(out['nccl_losses'])
out['loss_history_rel_diff_p2p_vs_nccl'] = float(np.max(np.abs(lp - ln) / np.abs(ln)))
out['obj_rel_diff_p2p_vs_nccl'] = float((res['p2p']['obj'] - res['nccl']['obj']).double().norm() / res['nccl']['obj'].double().norm())
out['world'] = world
l1 = np.array(out['nccl1_losses'])
out['loss_history_rel_diff_p2p_vs_nccl1'] = float(np.max(np.abs(lp - l1) / np.abs(l1)))
if rank == 0:
    print(json.dumps(out))
torch.distributed.destroy_process_group()
