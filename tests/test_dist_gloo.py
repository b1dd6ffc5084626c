"""CPU, world_size = 2 over gloo: the host-side data-parallel logic that replaces the reference's mpi4py calls
(SURVEY.md 8e): identical shuffled schedule on every rank, rank r takes entries [r*mb, (r+1)*mb) of each global chunk
(adorym/ptychography.py:846,905-907), gradients are summed -- NOT averaged -- (ptychography.py:1113-1114,
optimizers.py:1025-1079), objects are broadcast from rank 0 (ptychography.py:664-665)."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from adorym_b200 import dist as adist
    r, w, _ = adist.init_from_env('gloo')
    assert (r, w) == (rank, world)
    comm = adist.Comm()
    assert comm.Get_rank() == rank and comm.Get_size() == world
    res = {}
    # SUM all-reduce of a tensor gradient and of a numpy parameter gradient (no division by n_ranks)
    g = torch.full((4, 3, 2), float(rank + 1))
    res['allreduce_tensor'] = comm.allreduce(g).clone().numpy()
    res['allreduce_numpy'] = np.asarray(comm.allreduce(np.arange(5, dtype='float32') * (rank + 1)))
    # rank 0's object wins a bcast
    res['bcast'] = comm.bcast({'seed': 1234 + rank, 'probe': np.ones(3) * rank}, root=0)
    buf = np.full(6, rank, dtype='int64')
    comm.Bcast(buf, root=0)
    res['Bcast'] = buf
    # schedule: same on all ranks, disjoint per-rank slices of one angle each
    n_theta, n_pos, mb = 5, 12, 3
    chunks = adist.epoch_task_list(0, n_theta, n_pos, mb, world)
    res['chunks'] = np.concatenate(chunks)
    mine = [adist.rank_minibatch(chunks, i, rank, mb, world) for i in range(len(chunks))]
    res['mine'] = [(t, ind.tolist()) for t, ind in mine]
    comm.Barrier()
    out[rank] = res
    dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_logic():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    for r in (a, b):
        np.testing.assert_array_equal(r['allreduce_tensor'], np.full((4, 3, 2), 3.0))          # 1 + 2, not the mean
        np.testing.assert_array_equal(r['allreduce_numpy'], np.arange(5, dtype='float32') * 3)
        assert r['bcast']['seed'] == 1234 and np.all(r['bcast']['probe'] == 0)
        np.testing.assert_array_equal(r['Bcast'], np.zeros(6, dtype='int64'))
    np.testing.assert_array_equal(a['chunks'], b['chunks'])                                    # identical shuffle
    # every (angle, spot) pair of the epoch is visited exactly once across the two ranks
    seen = set()
    n_theta, n_pos, mb = 5, 12, 3
    for (ta, ia), (tb, ib) in zip(a['mine'], b['mine']):
        assert len(ia) == mb and len(ib) == mb and ia == sorted(ia) and ib == sorted(ib)       # sorted (quirk 7)
        for t, ind in ((ta, ia), (tb, ib)):
            for i in ind:
                assert (t, i) not in seen
                seen.add((t, i))
    assert len(seen) == n_theta * n_pos
