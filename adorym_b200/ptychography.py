"""reconstruct_ptychography(): the reference's driver signature (adorym/ptychography.py:54-174) thinly re-hosted
around the fused B200 path.  Data parallelism only (distribution_mode=None): one process per GPU, the minibatch
schedule of the reference (ptychography.py:786-850,901-909), SUM all-reduce of the object / parameter gradients
over NCCL (replacing mpi4py, :1113-1114), identical Adam update on every rank.

Out of scope (raise NotImplementedError when requested): shared_file / distributed_object modes, ePIE, multiscale,
CTF forward algorithm, optimizers other than adam / gd / momentum, probe-defocus / exit-wave-offset / slice-position
refinement.
"""
import datetime
import os
import time
import numpy as np
import torch

from . import ops
from .array_ops import ObjectFunction, Gradient, Mask
from .differentiator import Differentiator
from .dist import Comm, epoch_task_list, rank_minibatch
from .forward_model import (ForwardModel, PtychographyModel, SingleBatchFullfieldModel, SingleBatchPtychographyModel,
                            SparseMultisliceModel, MultiDistModel)
from .optimizers import (Optimizer, AdamOptimizer, GDOptimizer, MomentumOptimizer,
                         create_and_initialize_parameter_optimizers,
                         initialize_parameter_gradients, update_parameter_gradients, update_parameters)
from .propagate import get_kernel
from .regularizers import L1Regularizer, ReweightedL1Regularizer, TVRegularizer
from .util import initialize_probe, save_rotation_lookup, apply_rotation, RotationTable

PI = 3.1415927   # default of theta_end only (adorym/ptychography.py:52, quirk 11)


class _Dataset(object):
    """Measured data handle: `exchange/data` [n_theta, n_spots, py, px] + optional metadata."""

    def __init__(self, fname, save_path='.'):
        self.meta = {}
        if isinstance(fname, (np.ndarray, torch.Tensor)):
            self.data = fname
        else:
            path = os.path.join(save_path, fname)
            if path.endswith('.npy'):
                self.data = np.load(path, mmap_mode='r')
            elif path.endswith('.npz'):
                z = np.load(path)
                self.data = z['exchange/data'] if 'exchange/data' in z.files else z['data']
                self.meta = {k.split('/')[-1]: z[k] for k in z.files if k.startswith('metadata/')}
            else:
                try:
                    import h5py as h5
                except ImportError:
                    from . import h5lite as h5      # contiguous exchange/data through np.memmap, no h5py needed
                f = h5.File(path, 'r')
                self.data = f['exchange/data']
                if 'metadata' in f:
                    self.meta = {k: f['metadata'][k][...] for k in f['metadata'].keys()}
        self.shape = tuple(self.data.shape)


def _print(msg, rank, verbose=True):
    if verbose and rank == 0:
        print('[adorym_b200 %s] %s' % (datetime.datetime.now().strftime('%H:%M:%S'), msg), flush=True)


def reconstruct_ptychography(
        fname, obj_size, probe_pos=None, theta_st=0, theta_end=PI, n_theta=None, theta_downsample=None,
        energy_ev=None, psize_cm=None, free_prop_cm=None,
        raw_data_type='magnitude', is_minus_logged=False, slice_pos_cm_ls=None,
        n_epochs='auto', crit_conv_rate=0.03, max_nepochs=200,
        regularizers=None, alpha_d=None, alpha_b=None, gamma=1e-6,
        minibatch_size=None, multiscale_level=1, n_epoch_final_pass=None,
        initial_guess=None, random_guess_means_sigmas=(8.7e-7, 5.1e-8, 1e-7, 1e-8),
        n_batch_per_update=1, reweighted_l1=False, interpolation='bilinear',
        update_scheme='immediate', unknown_type='delta_beta', randomize_probe_pos=False,
        common_probe_pos=True, fix_object=False,
        optimize_object=True, optimizer='adam', learning_rate=1e-5, update_using_external_algorithm=None,
        optimizer_batch_number_increment='angle',
        finite_support_mask_path=None, shrink_cycle=None, shrink_threshold=1e-9,
        object_type='normal', non_negativity=False,
        forward_model='auto', forward_algorithm='fresnel', ctf_lg_kappa=1.7,
        binning=1, fresnel_approx=True, pure_projection=False, two_d_mode=False,
        probe_type='gaussian', probe_initial=None, probe_extra_defocus_cm=None, n_probe_modes=1,
        shared_probe_among_angles=True, rescale_probe_intensity=False,
        loss_function_type='lsq', poisson_multiplier=1., beamstop=None, normalize_fft=False, safe_zone_width=0,
        scale_ri_by_k=True, sign_convention=1, fourier_disparity=False,
        save_path='.', output_folder=None, save_intermediate=False, save_intermediate_level='batch', save_history=False,
        store_checkpoint=True, use_checkpoint=True, force_to_use_checkpoint=False, n_batch_per_checkpoint=10,
        save_stdout=False,
        cpu_only=False, core_parallelization=True, gpu_index=0, n_dp_batch=20,
        distribution_mode=None, dist_mode_n_batch_per_update=None, precalculate_rotation_coords=True,
        cache_dtype='float32', rotate_out_of_loop=False, n_split_mpi_ata='auto',
        optimize_probe=False, probe_learning_rate=1e-5, optimizer_probe=None,
        probe_update_delay=0, probe_update_limit=None,
        optimize_probe_defocusing=False, probe_defocusing_learning_rate=1e-5, optimizer_probe_defocusing=None,
        optimize_probe_pos_offset=False, probe_pos_offset_learning_rate=1e-2, optimizer_probe_pos_offset=None,
        optimize_prj_pos_offset=False, prj_pos_offset_learning_rate=1e-2, optimizer_prj_pos_offset=None,
        optimize_all_probe_pos=False, all_probe_pos_learning_rate=1e-2, optimizer_all_probe_pos=None,
        optimize_slice_pos=False, slice_pos_learning_rate=1e-4, optimizer_slice_pos=None,
        optimize_free_prop=False, free_prop_learning_rate=1e-2, optimizer_free_prop=None,
        optimize_prj_affine=False, prj_affine_learning_rate=1e-3, optimizer_prj_affine=None,
        optimize_tilt=False, tilt_learning_rate=1e-3, optimizer_tilt=None, initial_tilt=None,
        optimize_ctf_lg_kappa=False, ctf_lg_kappa_learning_rate=1e-3, optimizer_ctf_lg_kappa=None,
        other_params_update_delay=0,
        use_epie=False, epie_alpha=0.8,
        dynamic_rate=True, pupil_function=None, probe_circ_mask=0.9, dynamic_dropping=False, dropping_threshold=8e-5,
        backend='pytorch', debug=False, t_max_min=None, xpu=False, run_bfloat16=False, run_float64=False,
        **kwargs):
    """Same keyword surface as the reference.  Returns a dict with the final object, probe, refined parameters and
    the per-batch loss history (the reference returns None and writes TIFFs; files here are .npy)."""
    t_zero = time.time()
    # ---- options outside the accelerated path fail loudly ----------------------------------------------------
    if cpu_only:
        raise NotImplementedError('adorym_b200 has no CPU path (cpu_only=True); use the reference for CPU runs')
    if distribution_mode is not None:
        raise NotImplementedError('adorym_b200: distribution_mode=%r is replaced by NCCL data parallelism' % (distribution_mode,))
    for name, val in (('use_epie', use_epie), ('xpu', xpu),
                      ('run_bfloat16', run_bfloat16), ('run_float64', run_float64),
                      ('fourier_disparity', fourier_disparity), ('dynamic_dropping', dynamic_dropping)):
        if val:
            raise NotImplementedError('adorym_b200: %s is not on the accelerated path' % name)
    if multiscale_level != 1 or forward_algorithm != 'fresnel' or interpolation != 'bilinear':
        raise NotImplementedError('adorym_b200: multiscale_level / forward_algorithm / interpolation option')
    if backend not in ('pytorch', 'autograd'):
        raise ValueError('backend')
    if not shared_probe_among_angles or not common_probe_pos:
        raise NotImplementedError('adorym_b200: per-angle probes / per-angle probe positions')
    if update_using_external_algorithm is not None:
        raise NotImplementedError('adorym_b200: external update algorithms')

    comm = Comm()
    n_ranks, rank = comm.Get_size(), comm.Get_rank()
    if not torch.cuda.is_available():
        raise RuntimeError('adorym_b200 needs a CUDA device (B200); there is no CPU fallback')
    if n_ranks > 1:
        device_obj = torch.device('cuda', torch.cuda.current_device())
    else:
        device_obj = torch.device('cuda', gpu_index)
        torch.cuda.set_device(device_obj)
    verbose = kwargs.pop('verbose', True)
    max_batches = kwargs.pop('max_batches_for_test', None)    # testing hook: stop after this many minibatches
    # use_cuda_graph=True: the loss + gradient evaluation of a minibatch is captured once in a CUDA graph and replayed
    # (adorym_b200.graph): worth it for the launch-bound configurations (small minibatches, thin objects)
    use_cuda_graph = bool(kwargs.pop('use_cuda_graph', False))

    timestr = comm.bcast(datetime.datetime.today().strftime('%Y%m%d_%H%M%S') if rank == 0 else None, root=0)
    if output_folder is None:
        output_folder = 'recon_{}'.format(timestr)
    if save_path != '.' and not os.path.isabs(output_folder):
        output_folder = os.path.join(save_path, output_folder)
    stdout_options = {'save_stdout': save_stdout, 'output_folder': output_folder, 'timestamp': timestr}

    # ---- data + metadata (ptychography.py:237-323) -------------------------------------------------------------
    dset = _Dataset(fname, save_path)
    prj = dset.data
    obj_size = list(obj_size)
    if obj_size[-1] == 1:
        two_d_mode = True
    if n_theta is None:
        n_theta = prj.shape[0]
    if two_d_mode:
        n_theta = 1
    if 'theta' in dset.meta:
        theta_ls = np.asarray(dset.meta['theta'])
    else:
        theta_ls = np.linspace(theta_st, theta_end, n_theta, dtype='float32')
    if theta_downsample is not None:
        theta_ls = theta_ls[::theta_downsample]
        n_theta = len(theta_ls)
    if probe_pos is None:
        probe_pos = np.array(dset.meta['probe_pos_px']).astype(float)
    else:
        probe_pos = np.array(probe_pos).astype(float)
    if energy_ev is None:
        energy_ev = float(dset.meta['energy_ev'])
    if psize_cm is None:
        psize_cm = float(dset.meta['psize_cm'])
    if free_prop_cm is None:
        free_prop_cm = dset.meta['free_prop_cm']
    if np.array(free_prop_cm).size == 1:
        is_multi_dist = False
        if isinstance(free_prop_cm, np.ndarray):
            free_prop_cm = float(np.asarray(free_prop_cm).reshape(-1)[0])
    else:
        is_multi_dist = True
    if is_multi_dist:
        probe_size = list(obj_size[:2])
        subprobe_size = list(prj.shape[-2:])
    else:
        probe_size = list(prj.shape[-2:])
        subprobe_size = probe_size
    for kw in ['probe_size']:
        kwargs.pop(kw, None)
    if minibatch_size is None:
        minibatch_size = len(probe_pos)
    if minibatch_size > 1 and len(probe_pos) == 1:
        minibatch_size = 1      # ptychography.py:342-346
    ds_level = 1
    this_obj_size = obj_size
    is_sparse_multislice = slice_pos_cm_ls is not None and len(slice_pos_cm_ls) > 1      # ptychography.py:285-289
    if is_sparse_multislice:
        slice_pos_cm_ls = np.asarray(slice_pos_cm_ls, dtype='float64')
    if rank == 0:
        os.makedirs(output_folder, exist_ok=True)
    comm.Barrier()

    voxel_nm = np.array([psize_cm] * 3) * 1.e7 * ds_level
    lmbda_nm = 1240. / energy_ev
    delta_nm = voxel_nm[-1]
    h = get_kernel(delta_nm * binning, lmbda_nm, voxel_nm, probe_size, fresnel_approx=fresnel_approx,
                   sign_convention=sign_convention)
    if precalculate_rotation_coords and kwargs.pop('write_rotation_lookup_files', False):
        folder = 'arrsize_{}_{}_{}_ntheta_{}'.format(*this_obj_size, n_theta)
        if rank == 0 and not os.path.exists(folder):
            save_rotation_lookup(this_obj_size, theta_ls)
        comm.Barrier()

    # ---- object optimizer / object / forward model (ptychography.py:417-564) ---------------------------------------
    if isinstance(optimizer, Optimizer):
        opt = optimizer
        opt.name = 'obj'
    elif optimizer == 'adam':
        opt = AdamOptimizer('obj', output_folder=output_folder, options_dict={'step_size': learning_rate})
    elif optimizer == 'gd':
        opt = GDOptimizer('obj', output_folder=output_folder,
                          options_dict={'step_size': learning_rate, 'dynamic_rate': True, 'first_downrate_iteration': 20})
    elif optimizer == 'momentum':
        opt = MomentumOptimizer('obj', output_folder=output_folder, options_dict={'step_size': learning_rate})
    else:
        raise NotImplementedError('adorym_b200: optimizer=%r (adam, gd and momentum are on the accelerated path)' % (optimizer,))
    opt.create_container([*this_obj_size, 2], use_checkpoint, device_obj)
    opt.set_index_in_grad_return(0)
    opt_ls = [opt]

    obj = ObjectFunction([*this_obj_size, 2], output_folder=output_folder, ds_level=ds_level, object_type=object_type)
    seed = comm.bcast(int(time.time() / 60), root=0)
    obj.initialize_array(initial_guess=initial_guess, device=device_obj, random_guess_means_sigmas=random_guess_means_sigmas,
                         unknown_type=unknown_type, non_negativity=non_negativity, seed=seed)

    u_free = v_free = None
    forwardmodel_args = {'loss_function_type': loss_function_type, 'distribution_mode': distribution_mode,
                         'device': device_obj, 'common_vars_dict': locals(), 'raw_data_type': raw_data_type}
    if forward_model == 'auto':
        if is_multi_dist:
            forward_model = MultiDistModel(**forwardmodel_args)
        elif is_sparse_multislice:
            forward_model = SparseMultisliceModel(**forwardmodel_args)
        elif minibatch_size == 1 and len(probe_pos) == 1 and np.allclose(probe_pos[0], 0):
            forward_model = SingleBatchFullfieldModel(**forwardmodel_args)
        elif common_probe_pos and minibatch_size == 1 and len(probe_pos) > 1 and n_probe_modes == 1:
            forward_model = SingleBatchPtychographyModel(**forwardmodel_args)         # ptychography.py:540-541
        else:
            forward_model = PtychographyModel(**forwardmodel_args)
    else:
        forward_model = forward_model(**forwardmodel_args)
    _print('Forward model: %s' % type(forward_model).__name__, rank, verbose)

    if regularizers is None:
        regularizers = []
        if alpha_d not in [0, None]:
            if reweighted_l1:
                regularizers.append(ReweightedL1Regularizer(alpha_d, alpha_b, unknown_type=unknown_type))
            else:
                regularizers.append(L1Regularizer(alpha_d, alpha_b, unknown_type=unknown_type))
        if gamma not in [0, None]:
            regularizers.append(TVRegularizer(gamma, unknown_type=unknown_type))
    forward_model.add_regularizers(regularizers)
    reg_rwl1 = None
    reweighted_l1 = False           # quirk 15: the kwarg is overwritten by the presence of the instance
    for r in regularizers:
        if isinstance(r, ReweightedL1Regularizer):
            reg_rwl1 = r
            reweighted_l1 = True

    gradient = Gradient(obj, forward_model=forward_model)
    gradient.arr = torch.zeros([*this_obj_size, 2], dtype=torch.float32, device=device_obj)

    # N > 1 ranks on one NVSwitch node: the object and its gradient live in symmetric (peer-mapped) memory and
    # "all-reduce + Adam" (ptychography.py:1113-1125) is one fused kernel per rank (dist.PeerUpdate).  ADM_DP=nccl selects
    # the NCCL all-reduce (pipelined by y slabs) instead; so do the option combinations whose constraints are not fused.
    peer = None
    ops.OBJ_GRAD_OUT = None
    if (n_ranks > 1 and n_ranks <= 8 and os.environ.get('ADM_DP', 'p2p') == 'p2p' and isinstance(opt, AdamOptimizer)
            and unknown_type == 'delta_beta' and optimize_object and not fix_object and int(np.prod(this_obj_size)) % 2 == 0):
        from .dist import PeerUpdate
        peer = PeerUpdate([*this_obj_size, 2], device_obj)
        with torch.no_grad():
            peer.object_tensor().copy_(obj.arr.detach())
        obj.arr = peer.object_tensor().requires_grad_(True)
        if update_scheme == 'immediate' and not rotate_out_of_loop:
            ops.OBJ_GRAD_OUT = peer.gradient_tensor()        # the adjoint writes straight into the symmetric buffer
    # y slabs by which the gradient exchange (NCCL all-reduce, or the fused peer-memory update) is pipelined with the
    # transpose chain that still produces the gradient
    # (default: 4 slabs for NCCL; 1 = not pipelined for the peer-memory update, which measured faster that way: the exchange
    # kernel competes with the producers for the SMs and every slab costs a cross-GPU barrier -- DESIGN.md section 6)
    ops.GRAD_SLABS = int(os.environ.get('ADM_GRAD_SLABS', '1' if peer is not None else '4')) if n_ranks > 1 else 1

    mask = None
    if finite_support_mask_path is not None:
        mask = Mask(this_obj_size, finite_support_mask_path)
        if isinstance(finite_support_mask_path, (np.ndarray, torch.Tensor)):
            mask_arr = finite_support_mask_path
        elif str(finite_support_mask_path).endswith('.npy'):
            mask_arr = np.load(finite_support_mask_path)
        else:
            from PIL import Image, ImageSequence
            mask_arr = np.stack([np.array(fr) for fr in ImageSequence.Iterator(Image.open(finite_support_mask_path))])
        if tuple(np.asarray(mask_arr).shape) != tuple(this_obj_size):
            raise ValueError('finite-support mask has shape %s, the object is %s' % (tuple(np.asarray(mask_arr).shape),
                                                                                     tuple(this_obj_size)))
        mask.initialize_array_with_values(mask_arr, device=device_obj)

    # ---- probe (ptychography.py:607-667) ---------------------------------------------------------------------------
    if rank == 0:
        probe_init_kwargs = dict(kwargs)
        probe_init_kwargs.update(lmbda_nm=lmbda_nm, psize_cm=psize_cm)
        pr0, pi0 = initialize_probe(probe_size, probe_type, pupil_function=pupil_function, probe_initial=probe_initial,
                                    rescale_intensity=rescale_probe_intensity, prj=prj, raw_data_type=raw_data_type,
                                    normalize_fft=normalize_fft, n_probe_modes=n_probe_modes,
                                    sign_convention=sign_convention, extra_defocus_cm=probe_extra_defocus_cm,
                                    **probe_init_kwargs)
        pr0, pi0 = np.asarray(pr0), np.asarray(pi0)
        if pr0.ndim == 3:
            probe_real_np, probe_imag_np = pr0[:n_probe_modes], pi0[:n_probe_modes]
            if len(probe_real_np) != n_probe_modes:
                raise RuntimeError('Length of supplied supplied probe does not match number of probe modes.')
        elif n_probe_modes == 1:
            probe_real_np, probe_imag_np = pr0[None], pi0[None]
        else:
            # ptychography.py:644-646: modes drawn around the initial probe from the global numpy RNG
            rs = np.random.RandomState(seed)
            probe_real_np = np.stack([rs.normal(pr0, abs(pr0) * 0.2) for _ in range(n_probe_modes)])
            probe_imag_np = np.stack([rs.normal(pi0, abs(pi0) * 0.2) for _ in range(n_probe_modes)])
    else:
        probe_real_np = probe_imag_np = None
    probe_real_np = comm.bcast(probe_real_np, root=0)
    probe_imag_np = comm.bcast(probe_imag_np, root=0)
    probe_real = torch.tensor(probe_real_np, dtype=torch.float32, device=device_obj, requires_grad=True)
    probe_imag = torch.tensor(probe_imag_np, dtype=torch.float32, device=device_obj, requires_grad=True)

    # ---- refinable parameters (ptychography.py:673-735) ------------------------------------------------------------
    probe_pos_int = np.round(probe_pos).astype(int)
    if initial_tilt is None:                       # ptychography.py:680-684
        tilt_ls = np.zeros([3, n_theta])
        tilt_ls[0] = theta_ls
    else:
        tilt_ls = np.asarray(initial_tilt, dtype='float64')
    n_dists = len(free_prop_cm) if is_multi_dist else 1
    prj_affine_ls = np.tile(np.array([[1., 0, 0], [0, 1., 0]]).reshape([1, 2, 3]), [n_dists, 1, 1])
    optimizable_params = {'probe_real': probe_real, 'probe_imag': probe_imag,
                          'probe_defocus_mm': torch.tensor(0.0, device=device_obj),
                          'probe_pos_offset': torch.zeros([n_theta, 2], device=device_obj,
                                                          requires_grad=bool(optimize_probe_pos_offset)),
                          'prj_pos_offset': torch.zeros([n_theta, 2], device=device_obj,
                                                        requires_grad=bool(optimize_prj_pos_offset))}
    if optimize_tilt:
        # ptychography.py:725-726: the tilt table [3, n_theta] becomes a refinable tensor (row 0 = the rotation angles)
        optimizable_params['tilt_ls'] = torch.tensor(tilt_ls, dtype=torch.float32, device=device_obj, requires_grad=True)
    elif initial_tilt is not None:
        tilt_ls = torch.tensor(tilt_ls, dtype=torch.float32, device=device_obj)
    if is_multi_dist:
        ppc0 = np.zeros([n_dists, 2])
    else:
        ppc0 = np.tile(probe_pos - probe_pos_int, [n_theta, 1, 1])
    optimizable_params['probe_pos_correction'] = torch.tensor(ppc0, dtype=torch.float32, device=device_obj,
                                                              requires_grad=bool(optimize_all_probe_pos))
    optimizable_params['prj_affine_ls'] = torch.tensor(prj_affine_ls, dtype=torch.float32, device=device_obj,
                                                       requires_grad=bool(optimize_prj_affine))
    if is_sparse_multislice and optimize_slice_pos:
        # ptychography.py:719-720
        optimizable_params['slice_pos_cm_ls'] = torch.tensor(np.asarray(slice_pos_cm_ls, dtype='float64'), dtype=torch.float32,
                                                             device=device_obj, requires_grad=bool(optimize_slice_pos))
    if is_multi_dist and optimize_free_prop:
        # ptychography.py:709-710: the propagation distances become a refinable tensor
        optimizable_params['free_prop_cm'] = torch.tensor(np.asarray(free_prop_cm, dtype='float64'), dtype=torch.float32,
                                                          device=device_obj, requires_grad=True)
    opt_ls, opt_args_ls = create_and_initialize_parameter_optimizers(optimizable_params, locals())

    graphed = use_cuda_graph and n_ranks == 1
    if graphed:
        from .graph import GraphedDifferentiator
        if not (torch.is_tensor(forward_model.prj) and forward_model.prj.is_cuda):
            # the whole of exchange/data goes to the device when it comfortably fits (|.| as get_data applies it); otherwise
            # it stays on the host and every minibatch's frames are uploaded into the graph's static buffer
            nbytes = int(np.prod(prj.shape)) * 4
            if nbytes < 0.25 * torch.cuda.mem_get_info(device_obj)[0]:
                full = np.abs(np.asarray(prj[...])).astype('float32')
                forward_model.prj = torch.from_numpy(full).to(device_obj)
        prj = forward_model.prj
        diff = GraphedDifferentiator(forward_model)
    else:
        diff = Differentiator()
    diff.create_loss_node(forward_model.get_loss_function(), opt_args_ls)

    loss_history = []
    rot_table = None
    n_pos = len(probe_pos)
    cont = True
    i_epoch = 0
    if n_epochs == 'auto':
        n_epochs = max_nepochs
    while cont:
        t0 = time.time()
        n_tot_per_batch = minibatch_size * n_ranks
        ind_list_rand = epoch_task_list(i_epoch, n_theta, n_pos, minibatch_size, n_ranks, two_d_mode=two_d_mode,
                                        theta_ls=theta_ls, randomize_probe_pos=randomize_probe_pos,
                                        update_scheme=update_scheme)
        n_batch = len(ind_list_rand)
        i_opt_batch = 0     # sic: the reference re-derives it from the (zero) checkpoint offsets every epoch (:848)
        current_i_theta = -1
        initialize_gradients = True
        for i_batch in range(n_batch):
            t00 = time.time()
            if max_batches is not None and len(loss_history) >= max_batches:
                cont = False
                break
            if t_max_min is not None and (time.time() - t_zero) / 60 >= t_max_min:
                _print('Terminating because the time limit is reached.', rank, verbose)
                cont = False
                break
            this_i_theta, this_ind_batch = rank_minibatch(ind_list_rand, i_batch, rank, minibatch_size, n_ranks)
            this_pos_batch = probe_pos_int[this_ind_batch]
            is_last_batch_of_this_theta = i_batch == n_batch - 1 or ind_list_rand[i_batch + 1][0, 0] != this_i_theta
            if rotate_out_of_loop and not two_d_mode and this_i_theta != current_i_theta:
                # ptychography.py:918-943: the object is rotated ONCE when the angle changes; the minibatches of the angle
                # then work on obj.arr_rot (even after an immediate update of obj.arr: reference behaviour)
                if rot_table is None:
                    folder = 'arrsize_{}_{}_{}_ntheta_{}'.format(*this_obj_size, n_theta)
                    rot_table = RotationTable(this_obj_size, theta_ls, device_obj,
                                              src_folder=folder if precalculate_rotation_coords else None)
                with torch.no_grad():
                    obj.arr_rot = apply_rotation(obj.arr.detach(), rot_table.lut(this_i_theta))
                obj.arr_rot.requires_grad_(True)
            current_i_theta = this_i_theta

            if reweighted_l1:
                with torch.no_grad():
                    if i_batch % 10 == 0:
                        a = obj.arr.detach()
                        weight_l1 = a.max() / (a.abs() + 1e-4 * a.mean())
                    reg_rwl1.update_l1_weight(weight_l1)

            # ---- gradients (ptychography.py:1008-1042) ----
            local_vars = {'this_i_theta': this_i_theta, 'this_pos_batch': this_pos_batch, 'prj': prj,
                          'this_ind_batch': this_ind_batch, 'tilt_ls': tilt_ls, 'free_prop_cm': free_prop_cm,
                          'slice_pos_cm_ls': slice_pos_cm_ls,
                          'safe_zone_width': safe_zone_width, 'ctf_lg_kappa': ctf_lg_kappa}
            grad_func_args = {}
            for arg in forward_model.argument_ls:
                if arg == 'obj':
                    grad_func_args[arg] = obj.arr_rot if (rotate_out_of_loop and not two_d_mode) else obj.arr
                elif arg in optimizable_params:
                    grad_func_args[arg] = optimizable_params[arg]
                else:
                    grad_func_args[arg] = local_vars[arg]
            forward_model.update_loss_args(grad_func_args)
            grads = list(diff.get_gradients(**grad_func_args))
            if i_batch + 1 < n_batch:
                # measured frames of the next minibatch: host->device copy on a side stream under this step's kernels
                nxt_theta, nxt_ind = rank_minibatch(ind_list_rand, i_batch + 1, rank, minibatch_size, n_ranks)
                forward_model.prefetch_data(nxt_theta, forward_model.data_indices(nxt_ind),
                                            theta_downsample=forward_model.common_vars['theta_downsample'],
                                            ds_level=1 if is_multi_dist else forward_model.common_vars['ds_level'])

            if initialize_gradients:
                # ptychography.py:1065-1067: zeros + grads[0] -- the buffer becomes the new gradient (no fill, no add)
                # (graph mode: grads[0] is the graph's static output; it is accumulated into below only when the update is
                # deferred, and then it must not alias the buffer the next replay overwrites)
                gradient.arr = grads[0].clone() if (graphed and (update_scheme != 'immediate' or rotate_out_of_loop)) else grads[0]
                initialize_parameter_gradients(opt_ls, device_obj)
                initialize_gradients = False
            else:
                gradient.arr += grads[0]
            if rotate_out_of_loop and not two_d_mode:
                # ptychography.py:1067-1078: the (accumulated) gradient lives in the beam frame and is rotated back with the
                # inverse-angle lookup table -- a forward interpolation, not the transpose -- after EVERY minibatch
                # (reference behaviour, flagged by its own TODO at :1075)
                with torch.no_grad():
                    gradient.arr = apply_rotation(gradient.arr, rot_table.lut(this_i_theta, reverse=True))
            opt_ls = update_parameter_gradients(opt_ls, grads)

            if update_scheme == 'per angle' and not is_last_batch_of_this_theta:
                continue
            initialize_gradients = True

            # ---- all-reduce (SUM) + update (ptychography.py:1113-1163) ----
            slabs = ops.last_obj_grad_slabs
            pipelined = (n_ranks > 1 and peer is None and slabs is not None
                         and slabs[0].data_ptr() == gradient.arr.data_ptr() and isinstance(opt, AdamOptimizer)
                         and optimize_object and not fix_object)
            if n_ranks > 1 and not pipelined and peer is None:
                gradient.arr = comm.allreduce(gradient.arr)
            with torch.no_grad():
                if optimize_object and not fix_object:
                    extra = {}
                    fused_opt = isinstance(opt, (AdamOptimizer, GDOptimizer, MomentumOptimizer))
                    if fused_opt:
                        # constraints fused into the Adam epilogue (ptychography.py:1138-1142, array_ops.py:239-244)
                        extra['clip_nonneg'] = bool(non_negativity and unknown_type != 'real_imag')
                        if unknown_type == 'delta_beta' and object_type == 'absorption_only':
                            extra['zero_channel'] = 0
                        if unknown_type == 'delta_beta' and object_type == 'phase_only':
                            extra['zero_channel'] = 1
                        if mask is not None and unknown_type == 'delta_beta':
                            extra['mask'] = mask.mask
                    if peer is not None:
                        # pull-reduce over NVLink + Adam + push, one kernel (dist.PeerUpdate / adm_p2p_reduce_adam)
                        if gradient.arr.data_ptr() != peer.g.data_ptr():
                            peer.gradient_tensor().copy_(gradient.arr)
                        if (slabs is not None and slabs[0].data_ptr() == peer.g.data_ptr() and len(slabs[1]) > 1
                                and gradient.arr.data_ptr() == peer.g.data_ptr()):
                            new = peer.update_pipelined(opt, i_opt_batch, slabs[1], **opt.options_dict, **extra)
                        else:
                            new = peer.update(opt, i_opt_batch, **opt.options_dict, **extra)
                    elif pipelined:
                        # the all-reduce runs slab by slab under the adjoint kernels and under Adam (dist.py)
                        from .dist import allreduce_update_pipelined
                        new = allreduce_update_pipelined(opt, obj.arr, gradient, i_opt_batch, slabs[1], **opt.options_dict,
                                                         **extra)
                    else:
                        new = opt.apply_gradient(obj.arr, gradient, i_opt_batch, **opt.options_dict, **extra)
                    if not fused_opt:
                        if non_negativity and unknown_type != 'real_imag':
                            new = new.clamp_(min=0)
                        if unknown_type == 'delta_beta' and object_type == 'absorption_only':
                            new[..., 0] = 0
                        if unknown_type == 'delta_beta' and object_type == 'phase_only':
                            new[..., 1] = 0
                    if unknown_type == 'real_imag' and object_type in ('absorption_only', 'phase_only'):
                        nrm = torch.sqrt(new[..., 0] ** 2 + new[..., 1] ** 2)
                        if object_type == 'absorption_only':
                            new = torch.stack([nrm, torch.zeros_like(nrm)], -1)
                        else:
                            new = torch.stack([new[..., 0] / nrm, new[..., 1] / nrm], -1)
                    obj.arr = new.requires_grad_(True)
            optimizable_params = update_parameters(opt_ls, optimizable_params, locals())
            if mask is not None and not (isinstance(opt, (AdamOptimizer, GDOptimizer, MomentumOptimizer))
                                         and unknown_type == 'delta_beta' and optimize_object and not fix_object):
                obj.apply_finite_support_mask_to_array(mask, unknown_type=unknown_type, device=device_obj)
            if mask is not None and shrink_cycle is not None and i_batch % shrink_cycle == 0 and i_batch > 0:
                mask.update_mask_array(obj, shrink_threshold)

            current_loss = forward_model.current_loss         # host sync once per batch, as the reference (quirk 14)
            loss_history.append((i_epoch, i_batch, current_loss, time.time() - t_zero))
            _print('Epoch %d batch %d/%d angle %d: loss %.6e, %.3f s, %.1f patterns/s' % (
                i_epoch, i_batch, n_batch, this_i_theta, current_loss, time.time() - t00,
                minibatch_size * n_ranks / (time.time() - t00)), rank, verbose)
            if optimizer_batch_number_increment == 'angle':
                # ptychography.py:1267-1271 compares with the rank-local current_i_theta; ranks that hold different angles
                # of one chunk would then advance their Adam step index at different times and the replicas of the object
                # would drift apart (measured: 3e-5 absolute after 10 minibatches on 2 ranks).  The step index follows the
                # FIRST angle of the chunk on every rank: identical to the reference for one rank, consistent for N.
                chunk_theta = int(ind_list_rand[i_batch][0, 0])
                if i_batch == n_batch - 1 or ind_list_rand[i_batch + 1][0, 0] != chunk_theta:
                    i_opt_batch += 1
            elif optimizer_batch_number_increment == 'batch':
                i_opt_batch += 1

        if i_epoch == n_epochs - 1:
            cont = False
        _print('Epoch %d done in %.2f s' % (i_epoch, time.time() - t0), rank, verbose)
        i_epoch += 1
        if rank == 0 and kwargs.get('save_outputs', True):
            np.save(os.path.join(output_folder, 'obj_ds_1.npy'), obj.arr.detach().cpu().numpy())
            np.save(os.path.join(output_folder, 'probe_ds_1.npy'),
                    torch.stack([optimizable_params['probe_real'], optimizable_params['probe_imag']], -1).detach().cpu().numpy())
    comm.Barrier()
    ops.OBJ_GRAD_OUT = None
    extras = {k: optimizable_params[k].detach() for k in ('tilt_ls', 'prj_affine_ls', 'probe_pos_offset', 'prj_pos_offset', 'free_prop_cm', 'slice_pos_cm_ls')
              if k in optimizable_params}
    return {'obj': obj.arr.detach(), 'probe_real': optimizable_params['probe_real'].detach(),
            'probe_imag': optimizable_params['probe_imag'].detach(),
            'probe_pos_correction': optimizable_params['probe_pos_correction'].detach(), **extras,
            'loss_history': loss_history, 'output_folder': output_folder}
