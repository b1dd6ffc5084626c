"""CUDA-graph replay of the per-minibatch loss + gradient evaluation.

The reference spends its time between kernels: a minibatch of a small configuration (the CI shape, 2-D ptychography with
minibatches of a few dozen patterns, multi-distance holography) is ~30 launches driven by Python, 20 us apiece, for a few
us of GPU work each (these working sets are L2 resident and launch bound).  `GraphedDifferentiator` is a drop-in
for `Differentiator` (adorym/differentiator.py:21-42): the first call with a given signature runs the loss function and
`torch.autograd.grad` twice on a side stream (allocator warm-up), captures them once in a `torch.cuda.CUDAGraph`, and every
later call is ONE graph launch.  What changes from minibatch to minibatch never enters the captured launches as a host
value: the rotation grid, the patch positions, the gather indices into `exchange/data` and into the position-correction
table, and the angle index live in STATIC device buffers (`ForwardModel._static`) that are refilled -- one pinned-host
staging buffer, one asynchronous copy, plus one device-to-device copy of the angle's rotation grid -- before the replay.
Tensor arguments (object, probe, refinable parameters) are adopted in place: the fused optimizers update them in place, so
their storage -- and the graph's input pointers -- stay put; a tensor that arrives in new storage is copied into the
static one.

The returned gradients (and `forward_model.current_loss`) are the graph's static outputs: they are overwritten by the next
call, exactly like a gradient buffer that is reused; consume (or clone) them before calling again.

Measured data that live on the host (a pinned tensor, an HDF5 dataset) keep going through the model's own data path
(`get_data` / `prefetch_data`: pinned staging, side-stream copy) and land in a static device buffer before the replay.

Falls back to the eager path (with the reason in `.fallback_reason`) with a beamstop mask (its pixel count is read on the host), for sub-divided MultiDist frames, or when the capture fails.
"""
import numpy as np
import torch

from .differentiator import Differentiator


class _Entry(object):
    __slots__ = ('graph', 'static', 'S', 'stage', 'packed', 'layout', 'grads', 'has_grid', 'n_pos_ppc', 'host_frames')


class GraphedDifferentiator(Differentiator):

    def __init__(self, forward_model, max_graphs=16):
        super(GraphedDifferentiator, self).__init__()
        self.fm = forward_model
        self.max_graphs = max_graphs
        self._graphs = {}
        self._stream = None
        self.fallback_reason = None
        self.n_replays = 0

    # ---------------------------------------------------------------------------------------------------------
    def _unsupported(self, kwargs):
        fm = self.fm
        prj = getattr(fm, 'prj', None)
        if prj is None:
            return 'no measured data attached to the forward model'
        if fm.common_vars.get('beamstop', None) is not None:
            return 'beamstop mask'
        if fm.common_vars.get('optimize_slice_pos', False):
            return 'refinable slice positions (the transfer-function tables change with every update)'
        if 'safe_zone_width' in kwargs:
            n_d = len(kwargs['free_prop_cm'])
            if (kwargs.get('safe_zone_width') or 0) > 0 or prj.shape[1] // n_d != 1:
                return 'sub-divided MultiDist frames'
        return None

    def _key(self, kwargs):
        items = []
        for name, v in kwargs.items():
            if torch.is_tensor(v):
                items.append((name, tuple(v.shape), str(v.dtype), bool(v.requires_grad), str(v.device)))
            elif name in ('this_pos_batch', 'this_ind_batch'):
                items.append((name, len(v)))
            elif name == 'this_i_theta':
                items.append((name,))
            elif name == 'prj':
                items.append((name, id(v)))
            elif isinstance(v, np.ndarray):
                items.append((name, v.shape, hash(v.tobytes())))
            else:
                items.append((name, repr(v)))
        # the regulariser list and the reweighted-L1 weights are part of what the graph reads
        items.append(tuple((type(r).__name__, id(getattr(r, 'weight_l1', None))) for r in self.fm.reg_list))
        return tuple(items)

    def _fill(self, e, kwargs):
        """Refill the static device buffers for this minibatch (host index arithmetic + two copies)."""
        fm = self.fm
        idx = fm.static_indices(kwargs['this_i_theta'], kwargs['this_pos_batch'], kwargs['this_ind_batch'],
                                **({'n_pos_ppc': e.n_pos_ppc} if e.n_pos_ppc is not None else {}))
        # pinned staging buffers in rotation: the host may run several replays ahead of the device
        e.stage.append(e.stage.pop(0))
        host, ev = e.stage[0]
        ev.synchronize()
        for name, (a, b) in e.layout.items():
            host[a:b] = torch.from_numpy(np.ascontiguousarray(idx[name]).reshape(-1))
        e.packed.copy_(host, non_blocking=True)
        ev.record(torch.cuda.current_stream(e.packed.device))
        if e.has_grid:
            dev = e.packed.device
            e.S['grid'].copy_(fm._rot_table_for(dev).grid(int(kwargs['this_i_theta'])), non_blocking=True)
        if e.host_frames:
            # measured data resident on the host: this minibatch's frames through the model's own (prefetching) data path
            # into the static device buffer the captured launches read
            saved, fm._static = fm._static, None
            try:
                cv = fm.common_vars
                batch = fm.get_data(kwargs['this_i_theta'], fm.data_indices(kwargs['this_ind_batch']),
                                    theta_downsample=cv.get('theta_downsample', None),
                                    ds_level=cv.get('ds_level', 1) if 'free_prop_cm' not in kwargs else 1)
            finally:
                fm._static = saved
            e.S['frames'].copy_(batch, non_blocking=True)
        with torch.no_grad():
            for name, st in e.static.items():
                v = kwargs[name]
                if torch.is_tensor(st) and torch.is_tensor(v) and v.data_ptr() != st.data_ptr():
                    st.detach().copy_(v)

    def _build(self, key, kwargs):
        fm = self.fm
        dev = kwargs['obj'].device
        e = _Entry()
        e.host_frames = not (torch.is_tensor(fm.prj) and fm.prj.is_cuda)
        e.static = {}
        for name, v in kwargs.items():
            if torch.is_tensor(v) and v.is_cuda:
                e.static[name] = v.detach().requires_grad_(v.requires_grad)      # adopts the caller's storage
            else:
                e.static[name] = v
        ppc = kwargs.get('probe_pos_correction', None)
        e.n_pos_ppc = int(ppc.shape[1]) if (torch.is_tensor(ppc) and ppc.dim() == 3) else None
        idx = fm.static_indices(kwargs['this_i_theta'], kwargs['this_pos_batch'], kwargs['this_ind_batch'],
                                **({'n_pos_ppc': e.n_pos_ppc} if e.n_pos_ppc is not None else {}))
        e.layout, n = {}, 0
        for name, a in idx.items():
            e.layout[name] = (n, n + a.size)
            n += (a.size + 3) & ~3                    # 16-byte aligned segments
        e.stage = [(torch.zeros((n,), dtype=torch.int32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
        e.packed = torch.zeros((n,), dtype=torch.int32, device=dev)
        e.S = {name: e.packed[a:b] for name, (a, b) in e.layout.items()}
        e.S['pos'] = e.S['pos'].view(-1, 2)
        cv = fm.common_vars
        e.has_grid = not (cv.get('two_d_mode', False) or fm.rotate_out_of_loop or cv.get('optimize_tilt', False)
                          or cv.get('initial_tilt', None) is not None)
        e.S['grid'] = torch.empty_like(fm._rot_table_for(dev).grid(int(kwargs['this_i_theta']))) if e.has_grid else None
        e.S['frames'] = None
        if e.host_frames:
            cv = fm.common_vars
            b0 = fm.get_data(kwargs['this_i_theta'], fm.data_indices(kwargs['this_ind_batch']),
                             theta_downsample=cv.get('theta_downsample', None),
                             ds_level=cv.get('ds_level', 1) if 'free_prop_cm' not in kwargs else 1)
            e.S['frames'] = torch.empty_like(b0)
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        s = self._stream
        nodes_of = lambda kw: [v for i, v in enumerate(kw.values()) if i in self.opt_args_ls]
        fm._static = e.S
        try:
            self._fill(e, kwargs)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for _ in range(2):                    # warm-up on the capture stream: allocator pools, workspaces, plans
                    l = self.loss_object(**e.static)
                    torch.autograd.grad(l, nodes_of(e.static))
            s.synchronize()
            e.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(e.graph, stream=s, capture_error_mode='thread_local'):
                l = self.loss_object(**e.static)
                e.grads = torch.autograd.grad(l, nodes_of(e.static), retain_graph=False, create_graph=False, allow_unused=False)
            torch.cuda.current_stream(dev).wait_stream(s)
        finally:
            fm._static = None
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = e
        return e

    def get_gradients(self, **kwargs):
        """Same contract as Differentiator.get_gradients; see the module docstring for the lifetime of the result."""
        why = self._unsupported(kwargs)
        if why is not None:
            self.fallback_reason = why
            return super(GraphedDifferentiator, self).get_gradients(**kwargs)
        key = self._key(kwargs)
        e = self._graphs.get(key)
        if e is None:
            try:
                e = self._build(key, kwargs)
            except Exception as ex:                      # capture refused (an op that syncs, ...): stay eager
                self.fallback_reason = 'capture failed: %s' % (ex,)
                self.fm._static = None
                self._unsupported = lambda kw, _why=self.fallback_reason: _why
                torch.cuda.synchronize()
                return super(GraphedDifferentiator, self).get_gradients(**kwargs)
        self._fill(e, kwargs)
        e.graph.replay()
        self.n_replays += 1
        return e.grads
