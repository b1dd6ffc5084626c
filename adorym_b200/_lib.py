"""ctypes binding of the C ABI declared in include/adorym_b200.h (the hand-written sm_100a kernels).

There is NO CPU fallback: if the in-tree CUDA library is missing and cannot be built, or a tensor is not a
contiguous CUDA tensor, these functions raise.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libadorym_b200.so')

c_int, c_float, c_void_p, c_longlong = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong
c_double = ctypes.c_double


class MultisliceCfg(ctypes.Structure):
    """struct adm_multislice_cfg (include/adorym_b200.h)."""
    _fields_ = [('P', c_int), ('n_slices', c_int), ('B', c_int), ('M', c_int), ('Y', c_int), ('X', c_int),
                ('k1', c_float), ('sign_convention', c_int), ('unknown_type', c_int), ('free_prop', c_int),
                ('normalize_fft', c_int), ('probe_per_pattern', c_int), ('h_separable', c_int), ('h_per_slice', c_int),
                ('hfree_separable', c_int)]


P_CFG = ctypes.POINTER(MultisliceCfg)


class MultiDistCfg(ctypes.Structure):
    """struct adm_multidist_cfg (include/adorym_b200.h)."""
    _fields_ = [('N', c_int), ('M', c_int), ('D', c_int), ('k1', c_float), ('sign_convention', c_int),
                ('unknown_type', c_int)]


P_MDCFG = ctypes.POINTER(MultiDistCfg)

# name -> argtypes; every entry point returns int (0 = ok).  Must list every symbol of include/adorym_b200.h.
SIGNATURES = {
    'adm_abi_version': [],
    'adm_prepare_transfer_function': [c_int, c_void_p, c_int, c_float, c_void_p, c_void_p],
    'adm_prepare_transfer_function_separable': [c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p],
    'adm_transmission_fwd': [c_int, c_float, c_int, c_longlong, c_int, c_void_p, c_void_p, c_void_p],
    'adm_transmission_adj': [c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_multislice_workspace_bytes': [c_int],
    'adm_multislice_fwd': [P_CFG, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p],
    'adm_multislice_bwd': [P_CFG, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p],
    'adm_multislice_bwd_gpsi': [P_CFG, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p],
    'adm_detector_loss': [c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p, c_void_p, c_void_p],
    'adm_rotate_fwd': [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_rotate_adj': [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_adam_step': [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_double, c_double, c_double, c_double, c_int,
                      c_int, c_void_p, c_void_p],
    'adm_p2p_reduce_adam': [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_longlong, c_double, c_double, c_double,
                            c_double, c_int, c_int, c_void_p, c_void_p],
    'adm_sgd_step': [c_void_p, c_void_p, c_void_p, c_longlong, c_double, c_double, c_int, c_void_p, c_void_p],
    'adm_regularizer_l1tv': [c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p,
                             c_void_p],
    'adm_regularizer_l1tv_scaled': [c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p,
                                    c_void_p, c_int, c_void_p],
    'adm_regularizer_l1tv_range': [c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_int, c_void_p],
    'adm_object_to_transmission': [c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_void_p, c_void_p],
    'adm_transmission_to_object_adj': [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int,
                                       c_int, c_void_p, c_int, c_int, c_void_p],
    'adm_probe_shift_fwd': [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_probe_shift_bwd': [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_lf_supported': [c_int],
    'adm_lf_spectrum_order': [c_int, c_void_p],
    'adm_multidist_fwd': [P_MDCFG, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_multidist_bwd': [P_MDCFG, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                          c_void_p],
    'adm_frame_shift_fwd': [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_frame_shift_bwd': [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_detector_loss_large': [c_int, c_int, c_longlong, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_lf_patch_step_fwd': [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p],
    'adm_lf_patch_step_bwd': [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    'adm_lf_farfield_fwd': [c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p],
    'adm_lf_farfield_bwd': [c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p],
    'adm_frame_shift_complex_fwd': [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_frame_shift_complex_bwd': [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    'adm_affine_resample_fwd': [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    'adm_affine_resample_adj': [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    'adm_affine_resample_dmat': [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p],
}

RETURNS_LONGLONG = ('adm_multislice_workspace_bytes',)
ERRORS = {1: 'bad argument', 2: 'unsupported configuration', 3: 'CUDA launch error'}


def bind(cdll):
    """Attach argtypes/restype for every ABI symbol; raises AttributeError if one is missing."""
    for name, argtypes in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.argtypes = argtypes
        fn.restype = c_longlong if name in RETURNS_LONGLONG else c_int
    return cdll


_lib = None
launch_count = 0          # number of kernel-launching ABI calls made by this process (bench.py reports it)


def get_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _build   # nvcc is part of the image; this is a build step, not a fallback
            _build.build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('adorym_b200: CUDA library %s is missing; run `python -m adorym_b200.build`' % LIB_PATH)
        _lib = bind(ctypes.CDLL(LIB_PATH))
        if _lib.adm_abi_version() != 8:
            raise RuntimeError('adorym_b200: ABI version mismatch')
    return _lib


def check(rc, name):
    if rc != 0:
        if rc == 2:
            raise NotImplementedError('adorym_b200.%s: %s' % (name, ERRORS[2]))
        raise RuntimeError('adorym_b200.%s failed: %s' % (name, ERRORS.get(rc, rc)))


_timing = None           # None, or a list of (name, start_event, end_event) filled while timing is enabled


def enable_timing(on=True):
    """Record a CUDA-event pair around every ABI call on the current stream (bench.py's per-kernel roofline)."""
    global _timing
    _timing = [] if on else None


def timing_summary():
    """-> {name: (n_calls, total_ms)}; synchronises."""
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in (_timing or []):
        n, t = out.get(name, (0, 0.0))
        out[name] = (n + 1, t + e0.elapsed_time(e1))
    return out


_ptr_devices = set()     # devices of the tensors whose pointers were taken since the last call()


def call(name, *args):
    """Invoke one ABI entry point.  The kernels launch on the CURRENT device / stream of the calling thread, so every
    tensor handed over through ptr() since the previous call must live on that device: anything else would make the
    kernel dereference another GPU's memory on an unrelated stream."""
    global launch_count
    fn = getattr(get_lib(), name)
    if _ptr_devices:
        import torch
        cur = torch.cuda.current_device()
        bad = [d for d in _ptr_devices if d != cur]
        _ptr_devices.clear()
        if bad:
            raise RuntimeError('adorym_b200.%s: tensors live on cuda:%s but the current device is cuda:%d; wrap the call in '
                               '`with torch.cuda.device(tensor.device):` (or torch.cuda.set_device)' % (name, bad, cur))
    if _timing is not None:
        import torch
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _timing.append((name, e0, e1))
    else:
        rc = fn(*args)
    launch_count += 1
    check(rc, name)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('adorym_b200: expected a CUDA tensor (there is no CPU path)')
    if not t.is_contiguous():
        raise RuntimeError('adorym_b200: expected a contiguous tensor')
    _ptr_devices.add(t.device.index)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
