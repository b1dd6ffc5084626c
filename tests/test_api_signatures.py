"""CPU: the drop-in boundary.  (1) The public classes / functions keep the reference's Python signatures
-- checked against the argument lists written out below (with the reference file:line they come from) and, when the
reference tree is mounted (build container only), against the reference source itself parsed with `ast` (never
imported, never required).  (2) The C-ABI library is built in-tree, loads, and exports every symbol that
include/adorym_b200.h declares (no compute calls: there is no GPU here)."""
import ast
import inspect
import os
import re
import subprocess
import pytest

import adorym_b200 as ab
from adorym_b200 import _lib, build as abuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/adorym'

PREDICT_ARGS = ['obj', 'probe_real', 'probe_imag', 'probe_defocus_mm', 'probe_pos_offset', 'this_i_theta', 'this_pos_batch',
                'prj', 'probe_pos_correction', 'this_ind_batch', 'tilt_ls', 'prj_pos_offset']      # forward_model.py:179-181
MULTIDIST_ARGS = ['obj', 'probe_real', 'probe_imag', 'probe_defocus_mm', 'probe_pos_offset', 'this_i_theta',
                  'this_pos_batch', 'prj', 'probe_pos_correction', 'this_ind_batch', 'free_prop_cm', 'safe_zone_width',
                  'prj_affine_ls', 'ctf_lg_kappa', 'prj_pos_offset']                                  # forward_model.py:819-822
MSPB_ARGS = ['grid_batch', 'probe_real', 'probe_imag', 'energy_ev', 'psize_cm', 'delta_cm', 'free_prop_cm',
             'obj_batch_shape', 'kernel', 'fresnel_approx', 'pure_projection', 'binning', 'device', 'type',
             'normalize_fft', 'sign_convention', 'optimize_free_prop', 'u_free', 'v_free', 'scale_ri_by_k',
             'is_minus_logged', 'pure_projection_return_sqrt', 'kappa', 'repeating_slice', 'return_fft_time',
             'shift_exit_wave', 'return_intermediate_wavefields']                                     # propagate.py:131-137
ADAM_APPLY_ARGS = ['x', 'gradient', 'i_batch', 'step_size', 'b1', 'b2', 'eps']                       # optimizers.py:275
ADAM_DEFAULTS = {'step_size': 0.001, 'b1': 0.9, 'b2': 0.999, 'eps': 1e-7}


def _args(fn):
    a = inspect.getfullargspec(fn).args
    return a[1:] if a and a[0] == 'self' else a


def test_forward_model_signatures():
    for cls in (ab.PtychographyModel, ab.SingleBatchFullfieldModel):
        assert _args(cls.predict) == PREDICT_ARGS
        init = inspect.signature(cls.__init__).parameters
        for k in ('loss_function_type', 'distribution_mode', 'device', 'common_vars_dict', 'raw_data_type', 'simulation_mode'):
            assert k in init                                                                         # forward_model.py:41-42,166-167
    assert _args(ab.MultiDistModel.predict) == MULTIDIST_ARGS
    # argument_ls is what the driver builds its call from (ptychography.py:1017-1028)
    fm = ab.PtychographyModel.__new__(ab.PtychographyModel)
    ab.PtychographyModel.__init__(fm, common_vars_dict=None)
    assert fm.argument_ls == PREDICT_ARGS
    assert fm.get_argument_index('probe_real') == 1 and fm.get_argument_index('probe_pos_correction') == 8
    with pytest.raises(NotImplementedError):
        ab.PtychographyModel(distribution_mode='shared_file')


def test_operator_signatures():
    assert _args(ab.multislice_propagate_batch) == MSPB_ARGS
    assert _args(ab.apply_rotation) == ['obj', 'coord_old', 'interpolation', 'axis', 'device', 'reverse', 'override_backend']
    assert _args(ab.pad_object)[:4] == ['obj_rot', 'this_obj_size', 'probe_pos', 'probe_size']       # util.py:1327
    assert _args(ab.get_kernel) == ['dist_nm', 'lmbda_nm', 'voxel_nm', 'grid_shape', 'fresnel_approx', 'sign_convention']
    a = _args(ab.AdamOptimizer.apply_gradient)
    assert a[:7] == ADAM_APPLY_ARGS
    sig = inspect.signature(ab.AdamOptimizer.apply_gradient).parameters
    for k, v in ADAM_DEFAULTS.items():
        assert sig[k].default == v
    assert _args(ab.AdamOptimizer.__init__) == ['name', 'output_folder', 'distribution_mode', 'options_dict', 'forward_model']
    assert _args(ab.AdamOptimizer.create_container)[:3] == ['whole_object_size', 'use_checkpoint', 'device_obj']
    rp = inspect.signature(ab.reconstruct_ptychography).parameters
    for k in ('fname', 'obj_size', 'probe_pos', 'theta_st', 'theta_end', 'n_theta', 'energy_ev', 'psize_cm', 'free_prop_cm',
              'minibatch_size', 'n_epochs', 'alpha_d', 'alpha_b', 'gamma', 'learning_rate', 'probe_type',
              'optimize_probe', 'optimize_all_probe_pos', 'regularizers', 'optimizer', 'update_scheme', 'unknown_type',
              'loss_function_type', 'raw_data_type', 'n_probe_modes', 'binning', 'fresnel_approx', 'sign_convention'):
        assert k in rp, k                                                                            # ptychography.py:57-174


def _ref_args(path, cls, fn):
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    f = [n for n in body if isinstance(n, ast.FunctionDef) and n.name == fn][0]
    a = [x.arg for x in f.args.args]
    return a[1:] if a and a[0] == 'self' else a


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not mounted (only present in the build container)')
def test_signatures_against_reference_source():
    assert _ref_args(REF + '/forward_model.py', 'PtychographyModel', 'predict') == _args(ab.PtychographyModel.predict)
    assert _ref_args(REF + '/forward_model.py', 'SingleBatchFullfieldModel', 'predict') == _args(ab.SingleBatchFullfieldModel.predict)
    assert _ref_args(REF + '/forward_model.py', 'MultiDistModel', 'predict') == _args(ab.MultiDistModel.predict)
    assert _ref_args(REF + '/propagate.py', None, 'multislice_propagate_batch') == _args(ab.multislice_propagate_batch)
    assert _ref_args(REF + '/propagate.py', None, 'get_kernel') == _args(ab.get_kernel)
    assert _ref_args(REF + '/util.py', None, 'apply_rotation') == _args(ab.apply_rotation)
    assert _ref_args(REF + '/util.py', None, 'pad_object') == _args(ab.pad_object)
    ref = _ref_args(REF + '/optimizers.py', 'AdamOptimizer', 'apply_gradient')
    assert ref[:7] == _args(ab.AdamOptimizer.apply_gradient)[:7]
    assert _ref_args(REF + '/optimizers.py', 'GDOptimizer', 'apply_gradient')[:6] == _args(ab.GDOptimizer.apply_gradient)[:6]
    assert _ref_args(REF + '/optimizers.py', 'MomentumOptimizer', 'apply_gradient')[:5] == \
        _args(ab.MomentumOptimizer.apply_gradient)[:5]
    ours = set(inspect.signature(ab.reconstruct_ptychography).parameters)
    ref = _ref_args(REF + '/ptychography.py', None, 'reconstruct_ptychography')
    missing = [a for a in ref if a not in ours]
    assert not missing, 'reconstruct_ptychography lacks reference kwargs: %s' % missing


def test_c_abi_library_builds_loads_and_exports_header_symbols():
    lib = abuild.build()
    assert os.path.exists(lib)
    hdr = open(os.path.join(ROOT, 'include', 'adorym_b200.h')).read()
    declared = sorted(set(re.findall(r'^(?:int|long long)\s+(adm_\w+)\s*\(', hdr, flags=re.M)))
    assert len(declared) >= 14
    # every declared symbol has a ctypes signature and vice versa
    assert declared == sorted(_lib.SIGNATURES.keys())
    syms = subprocess.run(['nm', '-D', '--defined-only', lib], capture_output=True, text=True).stdout
    for s in declared:
        assert re.search(r'\sT\s+%s$' % s, syms, flags=re.M), 'not exported: ' + s
    # the product has no CPU path: CPU tensors are rejected before any call reaches the library
    import torch
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))
    # sm_100a only
    arch = subprocess.run(['cuobjdump', '--list-elf', lib], capture_output=True, text=True).stdout
    assert 'sm_100a' in arch and 'sm_90' not in arch
