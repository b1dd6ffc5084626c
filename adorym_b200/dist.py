"""One-process-per-GPU communicator with the (tiny) mpi4py surface the reference's hot loop uses
(adorym/pseudo.py:107-139 is the interface spec): Get_rank, Get_size, Barrier, bcast, allreduce(SUM).

Replaces the pickled, host-staged mpi4py collectives (adorym/ptychography.py:1113-1114, optimizers.py:1025-1079)
by torch.distributed: NCCL over NVLink 5 / NVSwitch for CUDA tensors, gloo for the CPU tests."""
import os
import numpy as np
import torch
import torch.distributed as dist


class Comm(object):
    def __init__(self):
        self._on = dist.is_available() and dist.is_initialized()

    def Get_rank(self):
        return dist.get_rank() if self._on else 0

    def Get_size(self):
        return dist.get_world_size() if self._on else 1

    def Barrier(self):
        if self._on:
            dist.barrier()

    def bcast(self, obj, root=0):
        if not self._on:
            return obj
        box = [obj]
        dist.broadcast_object_list(box, src=root)
        return box[0]

    def Bcast(self, arr, root=0):
        if self._on:
            t = torch.from_numpy(arr)
            if dist.get_backend() == 'nccl':
                tc = t.cuda()
                dist.broadcast(tc, src=root)
                t.copy_(tc.cpu())
            else:
                dist.broadcast(t, src=root)
        return arr

    def allreduce(self, x):
        """SUM all-reduce (no division: quirk 3).  CUDA tensors are reduced in place over NCCL."""
        if not self._on:
            return x
        if torch.is_tensor(x):
            dist.all_reduce(x, op=dist.ReduceOp.SUM)
            return x
        a = np.asarray(x)
        t = torch.from_numpy(np.ascontiguousarray(a))
        if dist.get_backend() == 'nccl':
            tc = t.cuda()
            dist.all_reduce(tc, op=dist.ReduceOp.SUM)
            return tc.cpu().numpy()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()


_comm_stream = {}


def allreduce_update_pipelined(opt, x, gradient, i_batch, slabs, **options):
    """SUM all-reduce of the object gradient + fused Adam, pipelined over the y slabs the backward pass produced
    (ops.last_obj_grad_slabs: [(y0, y1, event)]; the tomographic rotation is about y, so the transpose rotation, the
    regulariser and Adam are all independent per y plane).  Slab k is all-reduced on a side stream as soon as its event
    has fired -- while the adjoint kernels still produce slab k+1 -- and Adam runs on slab k while slab k+1 is on the
    wire.  Same sums and the same update as `gradient = comm.allreduce(gradient); opt.apply_gradient(...)`
    (adorym/ptychography.py:1113-1125): only the order in time changes.  Returns the updated x."""
    dev = x.device
    cs = _comm_stream.get(dev)
    if cs is None:
        cs = _comm_stream[dev] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    g = gradient.arr if hasattr(gradient, 'arr') else gradient
    if 'm' not in opt.params_whole_array_dict or opt.params_whole_array_dict['m'].shape != x.shape:
        opt.create_param_arrays(list(x.shape), device=dev)
    m, v = opt.params_whole_array_dict['m'], opt.params_whole_array_dict['v']
    mask = options.pop('mask', None)
    done = []
    for (y0, y1, ev) in slabs:
        cs.wait_event(ev)
        with torch.cuda.stream(cs):
            dist.all_reduce(g[y0:y1], op=dist.ReduceOp.SUM)
            e2 = torch.cuda.Event()
            e2.record(cs)
        done.append((y0, y1, e2))
    g.record_stream(cs)
    xd = x.detach()
    for (y0, y1, e2) in done:
        main.wait_event(e2)
        opt.apply_gradient(xd[y0:y1], g[y0:y1], i_batch, m=m[y0:y1], v=v[y0:y1], update_batch_count=False,
                           mask=None if mask is None else mask[y0:y1], **options)
    opt.i_batch += 1
    return xd


class PeerUpdate(object):
    """Data-parallel object update over NVLink peer memory for the ranks of one NVSwitch node (adm_p2p_reduce_adam).

    Owns two SYMMETRIC buffers (torch.distributed._symmetric_memory: every rank's copy is mapped into every process):
    the object `x` and the object gradient `g`.  The forward / adjoint kernels work on the local copies as usual
    (ops.OBJ_GRAD_OUT makes the adjoint write its gradient straight into `g`); `update()` then replaces
    `gradient = comm.allreduce(gradient); x = opt.apply_gradient(x, gradient, ...)` (adorym/ptychography.py:1113-1125) by
    barrier -> one fused kernel (pull my shard of every rank's g, sum in rank order, Adam on my shard, push the new shard
    into every rank's x) -> barrier.  Same SUM semantics, bit-identical replicas."""

    def __init__(self, shape, device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.shape = tuple(int(s) for s in shape)
        n = int(np.prod(self.shape))
        if n % 4:
            raise ValueError('PeerUpdate needs a volume whose element count is a multiple of 4')
        self.n = n
        self.x = symm.empty(n, dtype=torch.float32, device=device)
        self.g = symm.empty(n, dtype=torch.float32, device=device)
        self.hx = symm.rendezvous(self.x, group)
        self.hg = symm.rendezvous(self.g, group)
        self.world, self.rank = int(self.hx.world_size), int(self.hx.rank)
        if self.world > 8:
            raise NotImplementedError('PeerUpdate: at most 8 ranks (one NVSwitch node)')
        self._xp = (ctypes.c_void_p * self.world)(*[int(p) for p in self.hx.buffer_ptrs])
        self._gp = (ctypes.c_void_p * self.world)(*[int(p) for p in self.hg.buffer_ptrs])
        # NVLS multicast mappings (in-switch reduction / replication) when the fabric offers them.  Measured on B200: with 2
        # ranks plain peer loads / stores are faster (0.21 vs 0.35 ms: one peer, nothing for the switch to combine), with 8
        # the multicast path wins (0.30 vs 0.36 ms).  ADM_P2P_MULTICAST = 1 / 0 forces / disables it; default: >= 4 ranks.
        self.mc_x = self.mc_g = None
        mc = os.environ.get('ADM_P2P_MULTICAST', 'auto')
        want_mc = mc == '1' or (mc not in ('0', '1') and self.world >= 4)
        try:
            if want_mc and self.hx.has_multicast_support:
                mx, mg = int(self.hx.multicast_ptr), int(self.hg.multicast_ptr)
                if mx and mg:
                    self.mc_x, self.mc_g = ctypes.c_void_p(mx), ctypes.c_void_p(mg)
        except Exception:
            self.mc_x = self.mc_g = None
        self.x.zero_()
        self.g.zero_()

    def object_tensor(self):
        return self.x.view(self.shape)

    def gradient_tensor(self):
        return self.g.view(self.shape)

    def update(self, opt, i_batch, step_size=0.001, b1=0.9, b2=0.999, eps=1e-7, clip_nonneg=False, zero_channel=None,
               mask=None, **unused):
        """One data-parallel Adam step on the symmetric object from the symmetric gradients of all ranks."""
        import ctypes
        from . import ops
        shape = list(self.shape)
        if 'm' not in opt.params_whole_array_dict or list(opt.params_whole_array_dict['m'].shape) != shape:
            opt.create_param_arrays(shape, device=self.x.device)
        m, v = opt.params_whole_array_dict['m'], opt.params_whole_array_dict['v']
        if mask is not None and (mask.dtype != torch.float32 or not mask.is_contiguous() or mask.numel() * 2 != self.n):
            raise RuntimeError('PeerUpdate.update: mask must be a contiguous fp32 tensor with one element per voxel')
        flags = (1 if clip_nonneg else 0) | (2 if zero_channel == 0 else 0) | (4 if zero_channel == 1 else 0)
        self.hg.barrier(channel=0)          # every rank's gradient is complete (and its reads of x are done)
        ops.call('adm_p2p_reduce_adam', ctypes.cast(self._xp, ctypes.c_void_p), ctypes.cast(self._gp, ctypes.c_void_p),
                 self.mc_x, self.mc_g, self.world, self.rank, ops.ptr(m), ops.ptr(v), self.n, float(step_size), float(b1), float(b2), float(eps),
                 int(i_batch), flags, ops.ptr(mask), ops.stream_ptr())
        self.hx.barrier(channel=1)          # every rank's shard has landed in every copy of x
        opt.i_batch += 1
        return self.object_tensor()


    def update_pipelined(self, opt, i_batch, slabs, step_size=0.001, b1=0.9, b2=0.999, eps=1e-7, clip_nonneg=False,
                         zero_channel=None, mask=None, **unused):
        """The same update, pipelined by y slabs with the kernels that still PRODUCE the gradient: `slabs` =
        [(y0, y1, event)] as recorded by ops.MultisliceFn.backward (ops.GRAD_SLABS > 1; the rotation is about y, so the
        transpose chain finishes the gradient plane by plane).  On a side stream, per slab: wait for the slab's event ->
        cross-GPU barrier (every rank has finished this slab) -> the fused kernel on the slab (every rank reduces, updates
        and pushes ITS share of the slab), so the NVLink traffic of slab k runs under the production of slab k + 1 and only
        the last slab's exchange is exposed.  The object is last READ (regulariser gradient) before the first slab's event
        on every rank, so a peer's push into a finished slab cannot race with a reader."""
        import ctypes
        from . import ops
        shape = list(self.shape)
        fresh_state = False
        if 'm' not in opt.params_whole_array_dict or list(opt.params_whole_array_dict['m'].shape) != shape:
            opt.create_param_arrays(shape, device=self.x.device)
            fresh_state = True          # zero-filled on the CURRENT stream: the side stream must see that (first call only)
        m, v = opt.params_whole_array_dict['m'], opt.params_whole_array_dict['v']
        if mask is not None and (mask.dtype != torch.float32 or not mask.is_contiguous() or mask.numel() * 2 != self.n):
            raise RuntimeError('PeerUpdate.update_pipelined: mask must be a contiguous fp32 tensor with one element per voxel')
        flags = (1 if clip_nonneg else 0) | (2 if zero_channel == 0 else 0) | (4 if zero_channel == 1 else 0)
        row = self.n // self.shape[0]                       # elements per y plane
        if row % 4:
            return self.update(opt, i_batch, step_size, b1, b2, eps, clip_nonneg, zero_channel, mask)
        dev = self.x.device
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        if fresh_state or getattr(self, '_side_state_owner', None) is not opt:
            self._side.wait_stream(main)
            self._side_state_owner = opt
        mflat, vflat = m.view(-1), v.view(-1)
        kflat = mask.view(-1) if mask is not None else None
        for (y0, y1, ev) in slabs:
            off, cnt = y0 * row, (y1 - y0) * row
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                self.hg.barrier(channel=0)
                xp = (ctypes.c_void_p * self.world)(*[int(p) + 4 * off for p in self.hx.buffer_ptrs])
                gp = (ctypes.c_void_p * self.world)(*[int(p) + 4 * off for p in self.hg.buffer_ptrs])
                mcx = ctypes.c_void_p(self.mc_x.value + 4 * off) if self.mc_x is not None else None
                mcg = ctypes.c_void_p(self.mc_g.value + 4 * off) if self.mc_g is not None else None
                ops.call('adm_p2p_reduce_adam', ctypes.cast(xp, ctypes.c_void_p), ctypes.cast(gp, ctypes.c_void_p), mcx, mcg,
                         self.world, self.rank, ops.ptr(mflat[off:off + cnt]), ops.ptr(vflat[off:off + cnt]), cnt,
                         float(step_size), float(b1), float(b2), float(eps), int(i_batch), flags,
                         ops.ptr(kflat[off // 2:(off + cnt) // 2]) if kflat is not None else None, ops.stream_ptr())
        with torch.cuda.stream(self._side):
            self.hx.barrier(channel=1)
        main.wait_stream(self._side)
        opt.i_batch += 1
        return self.object_tensor()


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world_size, local_rank).  A no-op for single-process runs."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def rank_minibatch(ind_list_rand, i_batch, rank, minibatch_size, n_ranks):
    """This rank's (i_theta, sorted spot indices) of global chunk i_batch (adorym/ptychography.py:901-909);
    a short last chunk is topped up from chunk 0 (quirk 7)."""
    n_tot = minibatch_size * n_ranks
    c = ind_list_rand[i_batch]
    if len(c) < n_tot:
        c = np.concatenate([c, ind_list_rand[0][:n_tot - len(c)]])
        ind_list_rand[i_batch] = c
    this_i_theta = int(c[rank * minibatch_size, 0])
    this_ind_batch = np.sort(c[rank * minibatch_size:(rank + 1) * minibatch_size, 1])
    return this_i_theta, this_ind_batch


def epoch_task_list(i_epoch, n_theta, n_pos, minibatch_size, n_ranks, two_d_mode=False, theta_ls=None,
                    randomize_probe_pos=False, update_scheme='immediate'):
    """Shuffled (angle, spot) list of one epoch split into global chunks of minibatch_size * n_ranks
    (adorym/ptychography.py:786-847).  Deterministic in i_epoch, identical on every rank."""
    n_tot = minibatch_size * n_ranks
    rs = np.random.RandomState(i_epoch)
    if not two_d_mode:
        theta_ind_ls = np.arange(n_theta)
        rs.shuffle(theta_ind_ls)
    else:
        i0 = 0
        if theta_ls is not None:
            i0 = int(np.nonzero(np.abs(np.asarray(theta_ls) - theta_ls[0]) < 1e-5)[0][0])
        theta_ind_ls = np.array([i0])
    blocks = []
    for i_theta in theta_ind_ls:
        spots = np.arange(n_pos)
        if randomize_probe_pos:
            spots = rs.choice(spots, len(spots), replace=False)
        if update_scheme == 'immediate' and n_pos % minibatch_size != 0:
            spots = np.append(spots, rs.choice(spots[:-n_pos % minibatch_size], minibatch_size - (n_pos % minibatch_size),
                                               replace=False))
        elif update_scheme == 'per angle' and n_pos % n_tot != 0:
            spots = np.append(spots, rs.choice(spots[:-n_pos % n_tot], n_tot - (n_pos % n_tot), replace=False))
        blocks.append(np.stack([np.full(len(spots), i_theta), spots], axis=1))
    ind = np.concatenate(blocks, axis=0).astype('int32')
    return [ind[i:i + n_tot] for i in range(0, len(ind), n_tot)]
