"""ForwardModel classes with the reference's signatures (adorym/forward_model.py), backed by the fused kernels.

`predict` keeps the explicit positional argument list the driver builds by name from `argument_ls`
(adorym/ptychography.py:1017-1028, forward_model.py:173-175); `get_loss_function()` returns a function with that
same argument list whose result is a scalar torch tensor carrying a grad_fn, so `Differentiator.get_gradients`
(torch.autograd.grad over [obj, probe_real, probe_imag, ...]) works unchanged -- the graph consists of the
autograd.Functions of adorym_b200.ops instead of a taped slice loop.
"""
import inspect
import numpy as np
import torch

from . import ops
from .propagate import get_kernel, _free_prop_mode
from .util import RotationTable


def _np_int_positions(this_pos_batch):
    return np.round(np.asarray(this_pos_batch)).astype(int)


class ForwardModel(object):
    """Parent forward model (adorym/forward_model.py:28-161)."""

    def __init__(self, loss_function_type='lsq', distribution_mode=None, device=None, common_vars_dict=None,
                 raw_data_type='magnitude', simulation_mode=False):
        if distribution_mode is not None:
            raise NotImplementedError('adorym_b200: distribution_mode=%r is out of scope (data parallelism over NCCL '
                                      'replaces the MPI distribution modes)' % (distribution_mode,))
        self.loss_function_type = loss_function_type
        self.argument_ls = []
        self.regularizer_dict = {}
        self.distribution_mode = distribution_mode
        self.device = device
        self.simulation_mode = simulation_mode
        self._current_loss = 0
        self.raw_data_type = raw_data_type
        self.i_call = 0
        self.common_vars = common_vars_dict
        if common_vars_dict is not None:
            cv = common_vars_dict
            self.unknown_type = cv['unknown_type']
            self.normalize_fft = cv['normalize_fft']
            self.sign_convention = cv['sign_convention']
            self.rotate_out_of_loop = cv['rotate_out_of_loop']
            self.scale_ri_by_k = cv['scale_ri_by_k']
            self.is_minus_logged = cv['is_minus_logged']
            self.forward_algorithm = cv['forward_algorithm']
            self.stdout_options = cv.get('stdout_options', {})
            self.poisson_multiplier = cv['poisson_multiplier']
            self.common_probe_pos = cv['common_probe_pos']
            self.binning = cv['binning']
            self.prj = cv['prj']
        self.loss_args = {}
        self.reg_list = []
        self._plans = {}
        self._rot_table = None
        self.last_reg_value = None
        # CUDA-graph mode (adorym_b200.graph.GraphedDifferentiator): dict of STATIC device buffers that replace every
        # per-minibatch host value on the device side (rotation grid, patch positions, gather indices, angle index), so
        # that the captured launches can be replayed for any minibatch; None in eager mode
        self._static = None

    # -- the reference exposes current_loss as a float; reading it synchronises (forward_model.py:140-146) --
    @property
    def current_loss(self):
        v = self._current_loss
        return float(v.item()) if torch.is_tensor(v) else float(v)

    @current_loss.setter
    def current_loss(self, v):
        self._current_loss = v

    def update_loss_args(self, kwargs):
        self.loss_args = kwargs

    def add_regularizers(self, reg_list):
        self.reg_list = reg_list

    def get_regularization_value(self, obj, device=None):
        reg = None
        fused = self._fusable_regularizers()
        if fused is not None and obj.dim() == 4:
            # L1-type and / or TV terms on a delta_beta object: a single pass over the volume for all of them
            reg = ops.RegL1TVFn.apply(obj, *fused)
            self.last_reg_value = reg
            return reg
        for r in self.reg_list:
            v = r.get_value(obj, distribution_mode=self.distribution_mode, device=device)
            reg = v if reg is None else reg + v
        self.last_reg_value = reg
        return reg

    def _fusable_regularizers(self):
        """-> (weight_l1|None, alpha_d, alpha_b, gamma) when the regulariser list is any combination of at most one
        L1-type and at most one TV term on a delta_beta object (one fused kernel pass, ops.MultisliceFn `reg`), else
        None (the terms are then evaluated one by one through their get_value)."""
        from .regularizers import L1Regularizer, ReweightedL1Regularizer, TVRegularizer
        if len(self.reg_list) == 0 or len(self.reg_list) > 2:
            return None
        l1 = [r for r in self.reg_list if isinstance(r, (L1Regularizer, ReweightedL1Regularizer))]
        tv = [r for r in self.reg_list if isinstance(r, TVRegularizer)]
        if len(l1) + len(tv) != len(self.reg_list) or len(l1) > 1 or len(tv) > 1:
            return None
        if any(r.unknown_type != 'delta_beta' for r in self.reg_list) or self.unknown_type != 'delta_beta':
            return None
        w = None
        if l1 and isinstance(l1[0], ReweightedL1Regularizer):
            if l1[0].weight_l1 is None:
                return None
            w = l1[0].weight_l1
        return (w, l1[0].alpha_d if l1 else 0., l1[0].alpha_b if l1 else 0., tv[0].gamma if tv else 0.)

    def get_argument_index(self, arg):
        for i, a in enumerate(self.argument_ls):
            if a == arg:
                return i
        raise ValueError('{} is not in the argument list.'.format(arg))

    # ------------------------------------------------------------------------------------------------
    def _device(self, obj):
        if not obj.is_cuda:
            raise RuntimeError('adorym_b200 has no CPU path: the object must live on a CUDA device')
        return obj.device

    def _check_unsupported(self, names):
        cv = self.common_vars
        for n in names:
            if cv.get(n, False):
                raise NotImplementedError('adorym_b200: option %s is not on the accelerated path yet' % n)
        if cv.get('pure_projection', False) and (self.is_minus_logged or self.unknown_type != 'delta_beta'):
            raise NotImplementedError('adorym_b200: pure_projection is accelerated for delta_beta objects that are not '
                                      'minus-logged')
        if self.binning != 1 and self.unknown_type != 'delta_beta':
            raise NotImplementedError('adorym_b200: binning > 1 is accelerated for delta_beta objects only')
        if self.forward_algorithm != 'fresnel':
            raise NotImplementedError('adorym_b200: forward_algorithm=%r' % (self.forward_algorithm,))

    def data_indices(self, this_ind_batch):
        """Rows of exchange/data[i_theta] one minibatch reads (MultiDistModel: every distance's frame)."""
        return np.asarray(this_ind_batch)

    def static_indices(self, this_i_theta, this_pos_batch, this_ind_batch, n_pos_ppc=0):
        """Host side of graph mode: the int32 device-side indices one minibatch needs, as a dict of numpy arrays
        ('pos' [B,2] patch corners, 'ppc' flat rows of the position-correction table, 'data' flat rows of exchange/data,
        'theta' [1]).  Must mirror what predict / get_data index with in eager mode."""
        cv = self.common_vars
        it = int(this_i_theta)
        ind = np.asarray(this_ind_batch).astype('int64')
        td = cv.get('theta_downsample', None) or 1
        n_spots = int(self.prj.shape[1])
        return {'pos': _np_int_positions(this_pos_batch).astype('int32').reshape(-1, 2),
                'ppc': (it * int(n_pos_ppc) + ind).astype('int32'),
                'data': (it * td * n_spots + ind).astype('int32'),
                'theta': np.array([it], dtype='int32')}

    def _rotation_grid(self, this_i_theta, device):
        cv = self.common_vars
        if self._static is not None and self._static.get('grid') is not None:
            return self._static['grid']
        return self._rot_table_for(device).grid(this_i_theta)

    def _rot_table_for(self, device):
        cv = self.common_vars
        if self._rot_table is None:
            n_theta = cv['n_theta']
            folder = 'arrsize_{}_{}_{}_ntheta_{}'.format(*cv['this_obj_size'], n_theta)
            self._rot_table = RotationTable(cv['this_obj_size'], cv['theta_ls'], device,
                                            src_folder=folder if cv.get('precalculate_rotation_coords', True) else None)
        return self._rot_table

    def _object_grid(self, obj, this_i_theta):
        """Normalised rotation grid for ops.MultisliceFn (layout 'yxz'), or None when the object already is in the beam
        frame: two_d_mode, or rotate_out_of_loop (the driver rotates once per angle and hands over obj.arr_rot,
        ptychography.py:918-943,1011; forward_model.py:266)."""
        cv = self.common_vars
        dev = self._device(obj)
        if cv['two_d_mode'] or self.rotate_out_of_loop:
            return None
        return self._rotation_grid(this_i_theta, dev)

    def _large_frame_multislice(self, obj, this_i_theta, probe):
        """Full-field multislice on frames the shared-memory FFT cannot hold (128 .. 2048 pixels per side): the slice loop
        of propagate.py:208-258 with one fused large-frame step per slice -- psi <- IFFT2(H . FFT2(psi . T(slice))) is
        exactly the thin-object near-field kernel chain of ops.MultiDistFn (row pass: modulation + FFT, column pass:
        FFT . H . IFFT, row pass: IFFT) with the running wavefield in the role of the probe; autograd chains the per-slice
        analytic adjoints.  The last slice is not propagated inside the object (:253); a finite free_prop_cm is one more
        fused step with the free-space kernel (:271-280)."""
        cv = self.common_vars
        dev = self._device(obj)
        Y, X, Z = cv['this_obj_size']
        mode = _free_prop_mode(cv['free_prop_cm'])
        if mode == 'inf':
            raise NotImplementedError('adorym_b200: far-field detection of %d-pixel full-field frames is not accelerated' % Y)
        grid = self._object_grid(obj, this_i_theta)
        if grid is not None:
            rot = ops.RotateFn.apply(obj, grid)                                  # [Z, Y, X, 2]
        else:
            rot = obj.permute(2, 0, 1, 3).contiguous()
        key = ('lf', Y, mode, float(cv['free_prop_cm']) if mode == 'near' else 0., str(dev))
        plans = self._plans.get(key)
        if plans is None:
            psize_cm = cv['psize_cm'] * cv.get('ds_level', 1)
            voxel_nm = np.array([psize_cm] * 3) * 1.e7
            k1 = ops.k1_of(cv['energy_ev'], psize_cm, self.scale_ri_by_k)
            mk = lambda h: ops.MultiDistPlan(Y, [h], k1, dev, sign_convention=self.sign_convention, unknown_type=self.unknown_type)
            p_in = mk(cv['h']) if Z > 1 else None
            p_free = None
            if mode == 'near':
                p_free = mk(get_kernel(float(cv['free_prop_cm']) * 1e7, 1240. / cv['energy_ev'], voxel_nm, [Y, X],
                                       sign_convention=self.sign_convention))          # quirk 10: always paraxial
            plans = self._plans[key] = (p_in, p_free, k1)
        p_in, p_free, k1 = plans
        psi = probe
        for z in range(Z - 1):
            psi = ops.MultiDistFn.apply(rot[z], psi, p_in)[0]
        if p_free is not None:
            return ops.MultiDistFn.apply(rot[Z - 1], psi, p_free)                  # [1, M, N, N, 2]
        last = rot[Z - 1]
        if self.unknown_type == 'delta_beta':                                       # propagate.py:241
            e = torch.exp(-k1 * last[..., 1])
            ph = -self.sign_convention * k1 * last[..., 0]
            tr, ti = e * torch.cos(ph), e * torch.sin(ph)
        else:
            tr, ti = last[..., 0], last[..., 1]
        out = torch.stack([psi[..., 0] * tr - psi[..., 1] * ti, psi[..., 0] * ti + psi[..., 1] * tr], -1)
        return out.unsqueeze(0)

    def _exit_shift(self, prj_pos_offset, this_i_theta):
        """optimize_prj_pos_offset: this angle's refinable shift of the exit wave [1, 2] (forward_model.py:253-256), or
        None."""
        if not self.common_vars.get('optimize_prj_pos_offset', False) or prj_pos_offset is None:
            return None
        if self._static is not None:
            off = prj_pos_offset.index_select(0, self._static['theta'])
        else:
            off = prj_pos_offset[int(this_i_theta)]
        return off.reshape(1, 2).to(torch.float32)

    def _shift_exit_wave_and_detect(self, exit_wave, shift, plan, dev):
        """propagate.py:260-280 with shift_exit_wave: the exit wave of every pattern and mode is Fourier-shifted by the same
        refinable offset (util.py:380-397), THEN taken to the detector (far field / free-space step / nothing).  exit_wave
        [B, M, P, P, 2] comes from the multislice kernels run with free_prop = none; the detector transform is a second call
        of the same kernels on vacuum (one empty slice, every patch placed outside the object) with the shifted exit waves
        as per-pattern probes."""
        B, M, P = exit_wave.shape[0], exit_wave.shape[1], exit_wave.shape[2]
        shifted = ops.probe_shift(exit_wave.reshape(B * M, P, P, 2), shift)[0].reshape(B, M, P, P, 2)
        if plan.free_prop == 0:
            return shifted
        key = ('vac', id(plan))
        vplan = self._plans.get(key)
        if vplan is None:
            vplan = ops.MultislicePlan(P, 1, 1, 1, plan.k1, None, dev, sign_convention=plan.sign_convention,
                                       unknown_type='delta_beta', free_prop='none', normalize_fft=bool(plan.normalize_fft))
            vplan.free_prop, vplan.Hfp, vplan.Hfc = plan.free_prop, plan.Hfp, plan.Hfc
            self._plans[key] = vplan
        vac_obj = torch.zeros((1, 1, 1, 2), dtype=torch.float32, device=dev)
        vac_pos = torch.full((B, 2), -4 * P, dtype=torch.int32, device=dev)
        return ops.MultisliceFn.apply(vac_obj, None, shifted.contiguous(), vac_pos, vplan, None, 'yxz')

    def _multislice(self, obj, this_i_theta, probe, pos_t, plan, with_reg, in_beam_frame=False, exit_shift=None):
        """fused rotate + transmission + gather + multislice; with_reg: also the fused regulariser value (None when the
        regulariser list is not fusable) -> (det, reg_value | None, reg_was_fused).  in_beam_frame: the object has
        already been rotated (tilt refinement); the regulariser then acts on the UN-rotated object and is not fused."""
        grid = None if in_beam_frame else self._object_grid(obj, this_i_theta)
        if exit_shift is not None:
            if getattr(plan, 'large', False):
                raise NotImplementedError('adorym_b200: optimize_prj_pos_offset with probes larger than 72 pixels')
            # the slice loop ends at the exit wave; shift; detector transform
            key = ('exit', id(plan))
            eplan = self._plans.get(key)
            if eplan is None:
                import copy
                eplan = copy.copy(plan)
                eplan.free_prop = 0
                self._plans[key] = eplan
            out = self._multislice(obj, this_i_theta, probe, pos_t, eplan, with_reg, in_beam_frame)
            return (self._shift_exit_wave_and_detect(out[0], exit_shift, plan, obj.device),) + tuple(out[1:])
        if getattr(plan, 'large', False):
            # probes of 128 .. 2048 pixels: slice loop on the batched large-frame kernels; the regulariser is evaluated on
            # its own (ops.RegL1TVFn through get_regularization_value)
            return ops.MultisliceLargeFn.apply(obj, grid, probe, pos_t, plan), None, False
        reg = self._fusable_regularizers() if (with_reg and not in_beam_frame) else None
        sp = getattr(self, '_slice_pos_param', None)        # refinable slice positions (SparseMultisliceModel)
        if reg is not None:
            det, reg_val = ops.MultisliceFn.apply(obj, grid, probe, pos_t, plan, reg, 'yxz', sp)
            return det, reg_val, True
        return ops.MultisliceFn.apply(obj, grid, probe, pos_t, plan, None, 'yxz', sp), None, False

    def _plan(self, dev, P, Nz, Y, X, free_prop_cm):
        cv = self.common_vars
        mode = _free_prop_mode(free_prop_cm)
        # binning (propagate.py:207-254): Nz object slices -> ceil(Nz / binning) propagation steps; pure_projection
        # (propagate.py:158-186) is the single step that sums all slices, without any in-object propagation
        obj_slices = Nz
        binning = obj_slices if cv.get('pure_projection', False) else int(self.binning)
        Nz = -(-obj_slices // binning)
        key = (P, Nz, obj_slices, Y, X, mode, float(free_prop_cm) if mode == 'near' else 0., str(dev))
        plan = self._plans.get(key)
        if plan is None:
            ds_level = cv.get('ds_level', 1)
            psize_cm = cv['psize_cm'] * ds_level
            h_free = None
            if mode == 'near':
                voxel_nm = np.array([psize_cm] * 3) * 1.e7
                h_free = get_kernel(float(free_prop_cm) * 1e7, 1240. / cv['energy_ev'], voxel_nm, [P, P],
                                    sign_convention=self.sign_convention)   # quirk 10: always paraxial
            if P in ops.SUPPORTED_P:
                plan = ops.MultislicePlan(P, Nz, Y, X, ops.k1_of(cv['energy_ev'], psize_cm, self.scale_ri_by_k),
                                          cv['h'] if Nz > 1 else None, dev, sign_convention=self.sign_convention,
                                          unknown_type=self.unknown_type, free_prop=mode,
                                          normalize_fft=self.normalize_fft, h_free=h_free, binning=binning,
                                          obj_slices=obj_slices)
            else:
                plan = ops.LargeMultislicePlan(P, Nz, Y, X, ops.k1_of(cv['energy_ev'], psize_cm, self.scale_ri_by_k),
                                               cv['h'] if Nz > 1 else None, dev, sign_convention=self.sign_convention,
                                               unknown_type=self.unknown_type, free_prop=mode,
                                               normalize_fft=self.normalize_fft, h_free=h_free, binning=binning,
                                               obj_slices=obj_slices)
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------------------------------------
    def get_data(self, this_i_theta, this_ind_batch, theta_downsample=None, ds_level=1):
        """adorym/forward_model.py:113-119.  `prj` may be an HDF5-like dataset / ndarray (host; |.| and H2D copy per
        minibatch as in the reference) or a CUDA tensor of magnitudes already resident on the device."""
        if self._static is not None and self._static.get('frames') is not None:
            # graph mode with host-resident data: the minibatch's frames were uploaded into a static device buffer
            return self._static['frames']
        if theta_downsample is None:
            theta_downsample = 1
        prj = self.prj
        dev = self.device
        ind = np.asarray(this_ind_batch)
        it = int(this_i_theta) * theta_downsample
        pf = getattr(self, '_prefetched', None) if self._static is None else None
        if pf is not None:
            self._prefetched = None
            if pf[0] == (id(prj), it, ind.tobytes(), int(ds_level)):
                torch.cuda.current_stream(pf[1].device).wait_event(pf[2])
                pf[1].record_stream(torch.cuda.current_stream(pf[1].device))
                return pf[1]
        if torch.is_tensor(prj) and prj.is_cuda and self._static is not None:
            # graph mode: one gather over the flattened [n_theta * n_spots] frame table by a static device index
            frames = prj.reshape((-1,) + tuple(prj.shape[2:])).index_select(0, self._static['data'])
            batch = (frames.abs() if frames.is_complex() else frames).to(dev, dtype=torch.float32)
        elif torch.is_tensor(prj) and prj.is_cuda:
            frames = prj[it]
            if not (len(ind) == frames.shape[0] and np.array_equal(ind, np.arange(len(ind)))):
                frames = frames.index_select(0, torch.as_tensor(ind, device=prj.device, dtype=torch.long))
            batch = (frames.abs() if frames.is_complex() else frames).to(dev, dtype=torch.float32)
        elif torch.is_tensor(prj):
            # host tensor (ideally pinned): gather into a pinned staging buffer, asynchronous H2D copy
            frames = prj[it]
            if not (len(ind) == frames.shape[0] and np.array_equal(ind, np.arange(len(ind)))):
                shp = (len(ind),) + tuple(frames.shape[1:])
                if getattr(self, '_stage', None) is None or self._stage.shape != shp or self._stage.dtype != frames.dtype:
                    self._stage = torch.empty(shp, dtype=frames.dtype).pin_memory()
                ev = getattr(self, '_stage_event', None)
                if ev is not None:
                    ev.synchronize()          # the previous copy out of the staging buffer has finished
                torch.index_select(frames, 0, torch.as_tensor(ind, dtype=torch.long), out=self._stage)
                frames = self._stage
            batch = frames.to(dev, non_blocking=True)
            if frames is getattr(self, '_stage', None):
                self._stage_event = torch.cuda.Event()
                self._stage_event.record(torch.cuda.current_stream(dev))
            batch = (batch.abs() if batch.is_complex() else batch).to(torch.float32)
        else:
            a = np.abs(np.asarray(prj[it, ind]))
            batch = torch.from_numpy(np.ascontiguousarray(a, dtype='float32')).to(dev)
        if ds_level > 1:
            batch = batch[:, ::ds_level, ::ds_level]
        return batch

    def prefetch_data(self, this_i_theta, this_ind_batch, theta_downsample=None, ds_level=1):
        """Data path: start the host->device copy of the NEXT minibatch's measured frames on a side stream
        while the current minibatch computes; `get_data` with the same arguments then only waits for the copy's event.
        A no-op unless `prj` is a host tensor (the case that costs a synchronous copy inside the step)."""
        prj = self.prj
        if not torch.is_tensor(prj) or prj.is_cuda:
            return
        if theta_downsample is None:
            theta_downsample = 1
        dev = torch.device(self.device) if self.device is not None else torch.device('cuda', torch.cuda.current_device())
        ind = np.asarray(this_ind_batch)
        it = int(this_i_theta) * theta_downsample
        if getattr(self, '_pf_stream', None) is None:
            self._pf_stream = torch.cuda.Stream(device=dev)
        frames = prj[it]
        if not (len(ind) == frames.shape[0] and np.array_equal(ind, np.arange(len(ind)))):
            shp = (len(ind),) + tuple(frames.shape[1:])
            if getattr(self, '_pf_stage', None) is None or self._pf_stage.shape != shp or self._pf_stage.dtype != frames.dtype:
                self._pf_stage = torch.empty(shp, dtype=frames.dtype).pin_memory()
            elif getattr(self, '_pf_event', None) is not None:
                self._pf_event.synchronize()          # the previous copy out of the staging buffer has finished
            torch.index_select(frames, 0, torch.as_tensor(ind, dtype=torch.long), out=self._pf_stage)
            frames = self._pf_stage
        with torch.cuda.stream(self._pf_stream):
            batch = frames.to(dev, non_blocking=True)
            batch = (batch.abs() if batch.is_complex() else batch).to(torch.float32)
            if ds_level > 1:
                batch = batch[:, ::ds_level, ::ds_level]
            ev = torch.cuda.Event()
            ev.record(self._pf_stream)
        self._pf_event = ev
        self._prefetched = ((id(prj), it, ind.tobytes(), int(ds_level)), batch, ev)

    def _beamstop_mask(self, dev):
        bs = self.common_vars.get('beamstop', None)
        if bs is None:
            return None
        m = torch.as_tensor(bs, dtype=torch.float32, device=dev)
        return (m >= 1e-5).float().contiguous()

    def loss(self, this_pred_batch, this_prj_batch, obj):
        """Last-layer loss from predicted magnitudes (adorym/forward_model.py:121-147) -- API-compatible path built
        from library ops; `get_loss_function()` uses the fused kernel (`_fused_loss`) instead."""
        from .regularizers import mismatch_loss_from_magnitude
        mask = self._beamstop_mask(this_pred_batch.device)
        loss = mismatch_loss_from_magnitude(this_pred_batch, this_prj_batch, self.loss_function_type, self.raw_data_type,
                                            self.poisson_multiplier, mask)
        if len(self.reg_list) > 0:
            loss = loss + self.get_regularization_value(obj, device=self.device)
        self._current_loss = loss.detach()
        return loss

    def _fused_loss(self, det, this_prj_batch, obj, reg_val=None, reg_done=False):
        """detector magnitude + mismatch + residual in one kernel, then the regularisers (reg_done: their value came out
        of the fused object chain already)."""
        mask = self._beamstop_mask(det.device)
        loss, pred, lpp = ops.DetectorLossFn.apply(det, this_prj_batch, mask, self.loss_function_type,
                                                   self.raw_data_type, self.poisson_multiplier)
        self.last_pred = pred
        self.last_loss_per_pattern = lpp
        if reg_done:
            self.last_reg_value = reg_val
            loss = loss + reg_val
        elif len(self.reg_list) > 0:
            loss = loss + self.get_regularization_value(obj, device=self.device)
        self._current_loss = loss.detach()
        return loss

    def predict(self, *args, **kwargs):
        pass

    def get_loss_function(self):
        def calculate_loss(obj, this_i_theta, this_ind_batch, *args, **kwargs):
            theta_downsample = self.common_vars['theta_downsample']
            ds_level = self.common_vars['ds_level']
            this_pred_batch = self.predict(*args, **kwargs)
            this_prj_batch = self.get_data(this_i_theta, this_ind_batch, theta_downsample=theta_downsample, ds_level=ds_level)
            return self.loss(this_pred_batch, this_prj_batch, obj)
        return calculate_loss


class PtychographyModel(ForwardModel):
    """adorym/forward_model.py:164-401."""

    def __init__(self, loss_function_type='lsq', distribution_mode=None, device=None, common_vars_dict=None,
                 raw_data_type='magnitude', simulation_mode=False, run_bfloat16=False, run_float64=False):
        super(PtychographyModel, self).__init__(loss_function_type, distribution_mode, device, common_vars_dict,
                                                raw_data_type, simulation_mode=simulation_mode)
        if run_bfloat16 or run_float64:
            raise NotImplementedError('adorym_b200: the accelerated path computes in fp32 only')
        args = inspect.getfullargspec(self.predict).args
        args.pop(0)
        self.argument_ls = args
        self.run_bfloat16 = run_bfloat16
        self.run_float64 = run_float64

    def _detector_field(self, obj, probe_real, probe_imag, this_i_theta, this_pos_batch, probe_pos_correction,
                        this_ind_batch, with_reg=False, tilt_ls=None, probe_pos_offset=None, prj_pos_offset=None):
        """(shift probe) -> fused rotate + transmission + gather + multislice -> complex detector field [B, M, P, P, 2]
        (, regulariser value, fused flag)."""
        cv = self.common_vars
        # optimize_probe_defocusing: get_kernel() (numpy) on a tensor that requires grad fails in the reference's own
        # PyTorch backend (forward_model.py:247-251), so there is no behaviour to reproduce
        self._check_unsupported(['optimize_probe_defocusing'])
        dev = self._device(obj)
        probe_size = cv['probe_size']
        if probe_size[0] != probe_size[1]:
            raise NotImplementedError('adorym_b200: only square probes are supported')
        this_obj_size = cv['this_obj_size']
        pos = _np_int_positions(this_pos_batch)
        probe = torch.stack([probe_real, probe_imag], dim=-1)
        if probe.dim() == 3:
            probe = probe.unsqueeze(0)
        # (callers hand over tilt_ls / probe_pos_offset only where the reference's class of the same name uses them)
        if cv.get('optimize_probe_pos_offset', False) and probe_pos_offset is not None:
            # forward_model.py:258-260: every mode shifted by this angle's refinable offset (one Fourier shift)
            if self._static is not None:
                off = probe_pos_offset.index_select(0, self._static['theta'])
            else:
                off = probe_pos_offset[int(this_i_theta)]
            probe = ops.probe_shift(probe, off.reshape(1, 2).to(torch.float32))[0]
        tilted = tilt_ls is not None and (cv.get('optimize_tilt', False) or cv.get('initial_tilt', None) is not None)
        if tilted:
            # forward_model.py:262-276: three differentiable rotations (wrappers.py:1177-1217) replace the lookup table
            if cv['two_d_mode']:
                raise NotImplementedError('Tilt optimization is not impemented for distributed mode or when two_d_mode is on.')
            t = tilt_ls if torch.is_tensor(tilt_ls) else torch.as_tensor(np.asarray(tilt_ls), dtype=torch.float32, device=dev)
            t = t.to(dev)
            tcol = t.index_select(1, self._static['theta'])[:, 0] if self._static is not None else t[:, int(this_i_theta)]
            for axis in range(3):
                obj = ops.rotate_with_grad(obj, tcol[axis], axis)
        # per-position probe shift: quirk 8, positive corrections only trigger it (forward_model.py:297,304)
        if probe_pos_correction is not None and (cv['optimize_all_probe_pos']
                                                 or self._any_positive_correction(probe_pos_correction)):
            if self._static is not None:
                shift = probe_pos_correction.reshape(-1, 2).index_select(0, self._static['ppc']).to(torch.float32)
            else:
                idx = torch.as_tensor(np.asarray(this_ind_batch), device=dev, dtype=torch.long)
                shift = probe_pos_correction[int(this_i_theta)].index_select(0, idx).to(torch.float32)
            probe = ops.probe_shift(probe, shift)
        plan = self._plan(dev, probe_size[0], this_obj_size[2], this_obj_size[0], this_obj_size[1], cv['free_prop_cm'])
        # device copy of the integer positions, cached by content (the driver revisits the same minibatches every epoch)
        pos32 = np.ascontiguousarray(pos, dtype='int32')
        if self._static is not None:
            pos_t = self._static['pos']
        else:
            key = (pos32.tobytes(), str(dev))
            cache = self.__dict__.setdefault('_pos_cache', {})
            pos_t = cache.get(key)
            if pos_t is None:
                if len(cache) > 4096:
                    cache.clear()
                pos_t = cache[key] = torch.from_numpy(pos32).to(dev)
        out = self._multislice(obj, this_i_theta, probe, pos_t, plan, with_reg, in_beam_frame=tilted,
                               exit_shift=self._exit_shift(prj_pos_offset, this_i_theta))
        return out if with_reg else out[0]

    def _any_positive_correction(self, ppc):
        """forward_model.py:297,304: `any(probe_pos_correction > 1e-3)`.  The answer needs a device->host read; it is
        remembered for as long as the same tensor has not been written to (torch's version counter)."""
        key = (ppc.data_ptr(), ppc._version, tuple(ppc.shape))
        hit = getattr(self, '_ppc_check', None)
        if hit is None or hit[0] != key:
            hit = (key, bool((ppc > 1e-3).any()))
            self._ppc_check = hit
        return hit[1]

    def predict(self, obj, probe_real, probe_imag, probe_defocus_mm,
                probe_pos_offset, this_i_theta, this_pos_batch, prj,
                probe_pos_correction, this_ind_batch, tilt_ls, prj_pos_offset):
        """Detected magnitudes [minibatch_size, len_probe_y, len_probe_x] (adorym/forward_model.py:179-387)."""
        det = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch, probe_pos_correction,
                                   this_ind_batch, tilt_ls=tilt_ls, probe_pos_offset=probe_pos_offset,
                                   prj_pos_offset=prj_pos_offset)
        self.i_call += 1
        return ops.DetectorMagnitudeFn.apply(det)

    def get_loss_function(self):
        def calculate_loss(obj, probe_real, probe_imag, probe_defocus_mm,
                           probe_pos_offset, this_i_theta, this_pos_batch, prj,
                           probe_pos_correction, this_ind_batch, tilt_ls, prj_pos_offset):
            theta_downsample = self.common_vars['theta_downsample']
            ds_level = self.common_vars['ds_level']
            det, reg_val, reg_done = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch,
                                                          probe_pos_correction, this_ind_batch, with_reg=True,
                                                          tilt_ls=tilt_ls, probe_pos_offset=probe_pos_offset,
                                                          prj_pos_offset=prj_pos_offset)
            self.i_call += 1
            this_prj_batch = self.get_data(this_i_theta, this_ind_batch, theta_downsample=theta_downsample, ds_level=ds_level)
            return self._fused_loss(det, this_prj_batch, obj, reg_val, reg_done)
        return calculate_loss


class SingleBatchPtychographyModel(PtychographyModel):
    """adorym/forward_model.py:493-586: the reference's single-pattern fast path (it only avoids a stack / concatenate
    in Python); on the fused kernels a minibatch of one pattern is the general path with B = 1."""

    def __init__(self, loss_function_type='lsq', distribution_mode=None, device=None, common_vars_dict=None,
                 raw_data_type='magnitude', simulation_mode=False, *args, **kwargs):
        super(SingleBatchPtychographyModel, self).__init__(loss_function_type, distribution_mode, device, common_vars_dict,
                                                           raw_data_type, simulation_mode=simulation_mode, *args, **kwargs)

    def predict(self, obj, probe_real, probe_imag, probe_defocus_mm,
                probe_pos_offset, this_i_theta, this_pos_batch, prj,
                probe_pos_correction, this_ind_batch, tilt_ls, prj_pos_offset):
        if len(np.asarray(this_pos_batch)) != 1:
            raise ValueError('SingleBatchPtychographyModel takes one probe position per call')
        # forward_model.py:493-586 never shifts the probe (no probe_pos_correction branch)
        det = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch, None, this_ind_batch,
                                   prj_pos_offset=prj_pos_offset)
        self.i_call += 1
        return ops.DetectorMagnitudeFn.apply(det)

    def get_loss_function(self):
        def calculate_loss(obj, probe_real, probe_imag, probe_defocus_mm,
                           probe_pos_offset, this_i_theta, this_pos_batch, prj,
                           probe_pos_correction, this_ind_batch, tilt_ls, prj_pos_offset):
            theta_downsample = self.common_vars['theta_downsample']
            ds_level = self.common_vars['ds_level']
            det, reg_val, reg_done = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch, None,
                                                          this_ind_batch, with_reg=True, prj_pos_offset=prj_pos_offset)
            self.i_call += 1
            this_prj_batch = self.get_data(this_i_theta, this_ind_batch, theta_downsample=theta_downsample, ds_level=ds_level)
            return self._fused_loss(det, this_prj_batch, obj, reg_val, reg_done)
        return calculate_loss


class SparseMultisliceModel(PtychographyModel):
    """adorym/forward_model.py:589-806: a few thin slices at arbitrary positions `slice_pos_cm_ls`; the wave is propagated
    by the paraxial Fresnel kernel over each gap slice_pos[i+1] - slice_pos[i] (propagate.py:479-534,
    fresnel_propagate_wrapped :556-568), the transmission of every slice still uses the voxel size as its thickness
    (propagate.py:505).  Same fused kernels as PtychographyModel with one separable transfer function per gap
    (adm_multislice_cfg.h_per_slice).  Refinable slice positions (optimize_slice_pos): the tables are rebuilt from the
    current positions and dL/d slice_pos comes from the checkpointed forward / adjoint wavefields (ops._slice_pos_gradient)."""

    def __init__(self, loss_function_type='lsq', distribution_mode=None, device=None, common_vars_dict=None,
                 raw_data_type='magnitude', simulation_mode=False, *args, **kwargs):
        super(SparseMultisliceModel, self).__init__(loss_function_type, distribution_mode, device, common_vars_dict,
                                                    raw_data_type, simulation_mode=simulation_mode, *args, **kwargs)
        a = inspect.getfullargspec(self.predict).args
        a.pop(0)
        self.argument_ls = a
        self._sparse_plans = {}

    def _sparse_plan(self, dev, slice_pos_cm_ls):
        cv = self.common_vars
        # refinable slice positions (optimize_slice_pos): the tensor rides along to ops.MultisliceFn for its gradient; the
        # per-gap transfer functions are rebuilt from the current values (a device->host read of Nz numbers per update)
        self._slice_pos_param = slice_pos_cm_ls if (cv.get('optimize_slice_pos', False) and torch.is_tensor(slice_pos_cm_ls)
                                                    and slice_pos_cm_ls.requires_grad) else None
        pos = np.asarray(slice_pos_cm_ls.detach().cpu() if torch.is_tensor(slice_pos_cm_ls) else slice_pos_cm_ls, dtype='float64')
        P = cv['probe_size'][0]
        Y, X, Z = cv['this_obj_size']
        if len(pos) != Z:
            raise ValueError('slice_pos_cm_ls must hold one position per object slice')
        key = (tuple(pos.tolist()), str(dev), str(cv['free_prop_cm']))
        plan = self._sparse_plans.get(key)
        if plan is None:
            psize_cm = cv['psize_cm'] * cv.get('ds_level', 1)
            voxel_nm = np.array([psize_cm] * 3) * 1.e7
            lmbda_nm = 1240. / cv['energy_ev']
            gaps_nm = np.diff(pos) * 1e7
            # fresnel_propagate_wrapped -> get_kernel_wrapped with its default fresnel_approx=True (propagate.py:84-103)
            h_gaps = [get_kernel(float(d), lmbda_nm, voxel_nm, [P, P], sign_convention=self.sign_convention) for d in gaps_nm]
            mode = _free_prop_mode(cv['free_prop_cm'])
            h_free = None
            if mode == 'near':
                h_free = get_kernel(float(cv['free_prop_cm']) * 1e7, lmbda_nm, voxel_nm, [P, P],
                                    sign_convention=self.sign_convention)
            plan = ops.MultislicePlan(P, Z, Y, X, ops.k1_of(cv['energy_ev'], psize_cm, self.scale_ri_by_k), None, dev,
                                      sign_convention=self.sign_convention, unknown_type=self.unknown_type, free_prop=mode,
                                      normalize_fft=self.normalize_fft, h_free=h_free, h_gaps=h_gaps, lmbda_nm=lmbda_nm,
                                      voxel_nm=float(voxel_nm[0]))
            if len(self._sparse_plans) > 8:
                self._sparse_plans.clear()
            self._sparse_plans[key] = plan
        return plan

    def _plan(self, dev, P, Nz, Y, X, free_prop_cm):
        return self._sparse_plan(dev, self._slice_pos)

    def predict(self, obj, probe_real, probe_imag, probe_defocus_mm,
                probe_pos_offset, this_i_theta, this_pos_batch, prj,
                probe_pos_correction, this_ind_batch, slice_pos_cm_ls,
                prj_pos_offset):
        self._slice_pos = slice_pos_cm_ls
        det = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch, probe_pos_correction,
                                   this_ind_batch, probe_pos_offset=probe_pos_offset, prj_pos_offset=prj_pos_offset)
        self.i_call += 1
        return ops.DetectorMagnitudeFn.apply(det)

    def get_loss_function(self):
        def calculate_loss(obj, probe_real, probe_imag, probe_defocus_mm,
                           probe_pos_offset, this_i_theta, this_pos_batch, prj,
                           probe_pos_correction, this_ind_batch, slice_pos_cm_ls, prj_pos_offset):
            theta_downsample = self.common_vars['theta_downsample']
            ds_level = self.common_vars['ds_level']
            self._slice_pos = slice_pos_cm_ls
            det, reg_val, reg_done = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch,
                                                          probe_pos_correction, this_ind_batch, with_reg=True,
                                                          probe_pos_offset=probe_pos_offset, prj_pos_offset=prj_pos_offset)
            self.i_call += 1
            this_prj_batch = self.get_data(this_i_theta, this_ind_batch, theta_downsample=theta_downsample, ds_level=ds_level)
            return self._fused_loss(det, this_prj_batch, obj, reg_val, reg_done)
        return calculate_loss


class SingleBatchFullfieldModel(PtychographyModel):
    """adorym/forward_model.py:404-490: one full-field projection per step (probe covers the whole object)."""

    def __init__(self, loss_function_type='lsq', distribution_mode=None, device=None, common_vars_dict=None,
                 raw_data_type='magnitude', simulation_mode=False, *args, **kwargs):
        super(SingleBatchFullfieldModel, self).__init__(loss_function_type, distribution_mode, device, common_vars_dict,
                                                        raw_data_type, simulation_mode=simulation_mode, *args, **kwargs)

    def _detector_field(self, obj, probe_real, probe_imag, this_i_theta, this_pos_batch, probe_pos_correction,
                        this_ind_batch, with_reg=False, tilt_ls=None, probe_pos_offset=None, prj_pos_offset=None):
        # (forward_model.py:412-490 uses neither the tilt table nor the probe position offset)
        cv = self.common_vars
        dev = self._device(obj)
        probe_size = cv['probe_size']
        this_obj_size = cv['this_obj_size']
        if list(probe_size) != list(this_obj_size[:2]):
            raise ValueError('SingleBatchFullfieldModel needs probe_size == obj_size[:2]')
        probe = torch.stack([probe_real, probe_imag], dim=-1)
        if probe.dim() == 3:
            probe = probe.unsqueeze(0)
        if (probe_size[0] not in ops.SUPPORTED_P and probe_size[0] == probe_size[1] and ops.lf_supported(probe_size[0])
                and _free_prop_mode(cv['free_prop_cm']) != 'inf'):
            self._check_unsupported(['optimize_prj_pos_offset'])
            det = self._large_frame_multislice(obj, this_i_theta, probe.contiguous())
            return (det, None, False) if with_reg else det
        plan = self._plan(dev, probe_size[0], this_obj_size[2], this_obj_size[0], this_obj_size[1], cv['free_prop_cm'])
        pos_t = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        out = self._multislice(obj, this_i_theta, probe, pos_t, plan, with_reg,
                               exit_shift=self._exit_shift(prj_pos_offset, this_i_theta))
        return out if with_reg else out[0]

    def predict(self, obj, probe_real, probe_imag, probe_defocus_mm,
                probe_pos_offset, this_i_theta, this_pos_batch, prj,
                probe_pos_correction, this_ind_batch, tilt_ls, prj_pos_offset):
        det = self._detector_field(obj, probe_real, probe_imag, this_i_theta, this_pos_batch, probe_pos_correction,
                                   this_ind_batch, prj_pos_offset=prj_pos_offset)
        self.i_call += 1
        if self.simulation_mode:
            return det[:, 0, :, :, 0], det[:, 0, :, :, 1]      # (real, imag), forward_model.py:487-488
        return ops.DetectorMagnitudeFn.apply(det)


class MultiDistModel(ForwardModel):
    """adorym/forward_model.py:809-1092: one thin object, several near-field propagation distances.

    Accelerated for the single-block case (n_blocks == 1, safe_zone_width == 0, forward_algorithm == 'fresnel'):
      * frames of 128 .. 2048 pixels per side (power of two), one object slice: the large-frame kernels
        (ops.MultiDistFn: FFT2 of the modulated probe once for all distances, adjoint summed over distances in Fourier
        space), incl. the refinable per-distance shift of the measured frames (forward_model.py:1079-1088,
        ops.FrameShiftFn);
      * frames the shared-memory FFT holds (64, 72): the multislice kernels, one call per distance.
    Everything else raises NotImplementedError."""

    def __init__(self, loss_function_type='lsq', distribution_mode=None, device=None, common_vars_dict=None,
                 raw_data_type='magnitude', simulation_mode=False, *args, **kwargs):
        super(MultiDistModel, self).__init__(loss_function_type, distribution_mode, device, common_vars_dict,
                                             raw_data_type, simulation_mode=simulation_mode)
        a = inspect.getfullargspec(self.predict).args
        a.pop(0)
        self.argument_ls = a
        self._md_plans = {}

    def data_indices(self, this_ind_batch):
        cv = self.common_vars
        n_dists = len(np.asarray(cv['free_prop_cm']).reshape(-1))
        n_blocks = int(self.prj.shape[1]) // n_dists
        ind = np.asarray(this_ind_batch).astype('int64')
        return np.concatenate([ind + i * n_blocks for i in range(n_dists)])          # forward_model.py:1058-1061

    def static_indices(self, this_i_theta, this_pos_batch, this_ind_batch, n_pos_ppc=0):
        """Graph mode: exchange/data rows of every distance's frame (forward_model.py:1058-1061)."""
        out = super(MultiDistModel, self).static_indices(this_i_theta, this_pos_batch, this_ind_batch, n_pos_ppc)
        cv = self.common_vars
        n_dists = len(np.asarray(cv['free_prop_cm']).reshape(-1))
        n_spots = int(self.prj.shape[1])
        n_blocks = n_spots // n_dists
        ind = np.asarray(this_ind_batch).astype('int64')
        full = np.concatenate([ind + i * n_blocks for i in range(n_dists)])
        td = cv.get('theta_downsample', None) or 1
        out['data'] = (int(this_i_theta) * td * n_spots + full).astype('int32')
        return out

    def _md_plan(self, dev, N, dists):
        cv = self.common_vars
        key = (N, tuple(dists), str(dev))
        plan = self._md_plans.get(key)
        if plan is None:
            psize_cm = cv['psize_cm'] * cv.get('ds_level', 1)
            voxel_nm = np.array([psize_cm] * 3) * 1.e7
            # fresnel_propagate (propagate.py:537-553): always the paraxial kernel (quirk 10), built in float64
            hs = [get_kernel(d * 1e7, 1240. / cv['energy_ev'], voxel_nm, [N, N], sign_convention=self.sign_convention)
                  for d in dists]
            plan = ops.MultiDistPlan(N, hs, ops.k1_of(cv['energy_ev'], psize_cm, self.scale_ri_by_k), dev,
                                     sign_convention=self.sign_convention, unknown_type=self.unknown_type)
            self._md_plans[key] = plan
        return plan

    def _tiled_detector_fields(self, obj, probe_real, probe_imag, dists, szw, this_pos_batch):
        """Frames sub-divided into tiles (n_blocks > 1) and / or a safe zone (forward_model.py:917-1026): every tile of the
        minibatch reads the region [pos - szw, pos + subprobe + szw) of the object (vacuum outside) and of the probe ((1, 0)
        outside), is propagated on its own by the patch kernels (tile size 64 / 72: shared-memory multislice kernels, 128 ..
        2048: batched large-frame steps), and the safe zone is cropped from the detector field.  One probe mode: the
        reference indexes the imaginary probe patches by mode on the BATCH axis (:1005, quirk 9), i.e. every tile of a
        sub-batch of n_dp_batch tiles gets the imaginary patch of the sub-batch's first tile -- reproduced."""
        cv = self.common_vars
        dev = self._device(obj)
        Y, X, Z = cv['this_obj_size']
        sub = list(cv['subprobe_size'])
        if sub[0] != sub[1]:
            raise NotImplementedError('adorym_b200: only square tiles are supported')
        if Z != 1:
            # the in-object kernel `h` of the driver has the size of the whole frame (ptychography.py:389-392): the reference
            # cannot propagate a tile through more than one slice either
            raise NotImplementedError('adorym_b200: sub-divided MultiDistModel frames need a single object slice')
        if probe_real.dim() != 3 or probe_real.shape[0] != 1:
            raise NotImplementedError('adorym_b200: sub-divided MultiDistModel frames run with one probe mode (quirk 9)')
        tile = sub[0] + 2 * szw
        pos = _np_int_positions(this_pos_batch) - szw
        B = len(pos)
        # (1, 0)-padded probe, one [tile, tile] patch per tile
        lo_y, lo_x = max(0, -int(pos[:, 0].min())), max(0, -int(pos[:, 1].min()))
        hi_y, hi_x = max(0, int(pos[:, 0].max()) + tile - Y), max(0, int(pos[:, 1].max()) + tile - X)
        pad = [lo_x, hi_x, lo_y, hi_y]
        pr = torch.nn.functional.pad(probe_real, pad, value=1.)
        pi = torch.nn.functional.pad(probe_imag, pad, value=0.)
        n_dp = int(cv.get('n_dp_batch', 20))
        pr_ls, pi_ls = [], []
        for j in range(B):
            y0, x0 = int(pos[j, 0]) + lo_y, int(pos[j, 1]) + lo_x
            j0 = (j // n_dp) * n_dp                                      # first tile of this sub-batch (quirk 9)
            yq, xq = int(pos[j0, 0]) + lo_y, int(pos[j0, 1]) + lo_x
            pr_ls.append(pr[:, y0:y0 + tile, x0:x0 + tile])
            pi_ls.append(pi[:, yq:yq + tile, xq:xq + tile])
        probe = torch.stack([torch.stack(pr_ls), torch.stack(pi_ls)], dim=-1).contiguous()      # [B, 1, tile, tile, 2]
        pos_t = torch.from_numpy(np.ascontiguousarray(pos, dtype='int32')).to(dev)
        dets = []
        for d in dists:
            plan = self._plan(dev, tile, Z, Y, X, d)
            det = self._multislice(obj, 0, probe, pos_t, plan, False)[0]                         # [B, 1, tile, tile, 2]
            dets.append(det[:, :, szw:szw + sub[0], szw:szw + sub[1]] if szw > 0 else det)
        return torch.cat(dets, 0).contiguous()

    def _detector_fields(self, obj, probe_real, probe_imag, free_prop_cm, safe_zone_width, prj, this_pos_batch=None):
        cv = self.common_vars
        self._check_unsupported(['optimize_probe_defocusing', 'optimize_probe_pos_offset', 'optimize_prj_pos_offset',
                                 'optimize_ctf_lg_kappa'])
        refine_dist = bool(cv.get('optimize_free_prop', False))
        dists = [float(d) for d in np.asarray(free_prop_cm.detach().cpu() if torch.is_tensor(free_prop_cm) else free_prop_cm)]
        n_dists = len(dists)
        szw = int(safe_zone_width or 0)
        if szw > 0 or prj.shape[1] // n_dists != 1:
            if refine_dist:
                raise NotImplementedError('adorym_b200: optimize_free_prop with sub-divided frames / a safe zone')
            return self._tiled_detector_fields(obj, probe_real, probe_imag, dists, szw, this_pos_batch)
        dev = self._device(obj)
        this_obj_size = cv['this_obj_size']
        P = this_obj_size[0]
        if this_obj_size[0] != this_obj_size[1]:
            raise NotImplementedError('adorym_b200: only square frames are supported')
        probe = torch.stack([probe_real, probe_imag], dim=-1)
        if probe.dim() == 3:
            probe = probe.unsqueeze(0)
        if ops.lf_supported(P):
            if this_obj_size[2] != 1 or not cv['two_d_mode']:
                raise NotImplementedError('adorym_b200: large-frame MultiDistModel needs a single object slice (two_d_mode)')
            if any(d <= 0 for d in dists):
                raise NotImplementedError('adorym_b200: MultiDistModel needs positive propagation distances')
            if refine_dist:
                # refinable distances (optimize_free_prop): in-graph paraxial transfer functions + analytic dL/d dist
                psize_cm = cv['psize_cm'] * cv.get('ds_level', 1)
                geom = dict(N=P, k1=ops.k1_of(cv['energy_ev'], psize_cm, self.scale_ri_by_k),
                            sign_convention=self.sign_convention, unknown_type=self.unknown_type,
                            lmbda_nm=1240. / cv['energy_ev'], voxel_nm=psize_cm * 1.e7)
                fp = free_prop_cm if torch.is_tensor(free_prop_cm) else torch.tensor(dists, dtype=torch.float32, device=dev)
                return ops.MultiDistRefineFn.apply(obj.reshape(P, P, 2), probe, fp, geom)
            plan = self._md_plan(dev, P, dists)
            return ops.MultiDistFn.apply(obj.reshape(P, P, 2), probe, plan)      # [n_dists, M, P, P, 2]
        if refine_dist:
            raise NotImplementedError('adorym_b200: optimize_free_prop needs large-frame sizes (128 .. 2048 pixels)')
        pos_t = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        dets = []
        for d in dists:
            plan = self._plan(dev, P, this_obj_size[2], this_obj_size[0], this_obj_size[1], d)
            dets.append(self._multislice(obj, 0, probe, pos_t, plan, False)[0])
        return torch.cat(dets, 0)      # [n_dists, M, P, P, 2]

    def predict(self, obj, probe_real, probe_imag, probe_defocus_mm,
                probe_pos_offset, this_i_theta, this_pos_batch, prj,
                probe_pos_correction, this_ind_batch, free_prop_cm, safe_zone_width, prj_affine_ls, ctf_lg_kappa,
                prj_pos_offset):
        det = self._detector_fields(obj, probe_real, probe_imag, free_prop_cm, safe_zone_width, prj, this_pos_batch)
        self.i_call += 1
        return ops.DetectorMagnitudeFn.apply(det)

    def get_loss_function(self):
        def calculate_loss(obj, probe_real, probe_imag, probe_defocus_mm,
                           probe_pos_offset, this_i_theta, this_pos_batch, prj,
                           probe_pos_correction, this_ind_batch, free_prop_cm, safe_zone_width,
                           prj_affine_ls, ctf_lg_kappa, prj_pos_offset):
            theta_downsample = self.common_vars['theta_downsample']
            if theta_downsample is None:
                theta_downsample = 1
            det = self._detector_fields(obj, probe_real, probe_imag, free_prop_cm, safe_zone_width, prj, this_pos_batch)
            self.i_call += 1
            n_dists = len(free_prop_cm)
            n_blocks = prj.shape[1] // n_dists
            ind = np.asarray(this_ind_batch)
            full = np.concatenate([ind + i * n_blocks for i in range(n_dists)])       # forward_model.py:1058-1061
            this_prj_batch = self.get_data(this_i_theta, full, theta_downsample=theta_downsample, ds_level=1)
            large = ops.lf_supported(det.shape[2])
            if self.common_vars.get('optimize_prj_affine', False):
                # refinable affine alignment of every distance's hologram (forward_model.py:1066-1072,
                # wrappers.py:1158-1174): one resampling kernel, a 2x3 matrix per frame, d/d matrix by a reduction kernel
                if not large or self.common_vars.get('beamstop', None) is not None:
                    raise NotImplementedError('adorym_b200: affine refinement of the data needs large-frame sizes, no beamstop')
                nb = len(ind)
                aff = prj_affine_ls[:n_dists].to(torch.float32)
                mats = aff.repeat_interleave(nb, dim=0) if nb > 1 else aff
                this_prj_batch = ops.AffineResampleFn.apply(this_prj_batch, mats, (0, 1, 2))
            if self.common_vars.get('optimize_all_probe_pos', False):
                # per-distance sub-pixel shift of the measured frames, real part kept (forward_model.py:1079-1088)
                if not large:
                    raise NotImplementedError('adorym_b200: measured-frame shift refinement needs large-frame sizes')
                this_prj_batch = ops.FrameShiftFn.apply(this_prj_batch, probe_pos_correction[:n_dists])
            if large and self.common_vars.get('beamstop', None) is None:
                loss, pred, lpp = ops.DetectorLossLargeFn.apply(det, this_prj_batch, self.loss_function_type,
                                                                self.raw_data_type, self.poisson_multiplier)
                self.last_pred, self.last_loss_per_pattern = pred, lpp
                if len(self.reg_list) > 0:
                    loss = loss + self.get_regularization_value(obj, device=self.device)
                self._current_loss = loss.detach()
                return loss
            return self._fused_loss(det, this_prj_batch, obj)
        return calculate_loss
