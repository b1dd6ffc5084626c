"""adorym_b200 -- B200-native (sm_100a) implementation of Adorym's multislice forward-model + gradient + update
hot path, behind the reference's own Python API (reconstruct_ptychography, the ForwardModel classes,
AdamOptimizer).  See DESIGN.md / INTEGRATION.md.  There is no CPU fallback."""
from .propagate import multislice_propagate_batch, get_kernel, gen_freq_mesh
from .util import (apply_rotation, pad_object, calculate_pad_len, save_rotation_lookup, read_origin_coords,
                   initialize_probe, generate_gaussian_map)
from .forward_model import (ForwardModel, PtychographyModel, SingleBatchFullfieldModel, SingleBatchPtychographyModel,
                            SparseMultisliceModel, MultiDistModel)
from .optimizers import Optimizer, AdamOptimizer, GDOptimizer, MomentumOptimizer
from .regularizers import Regularizer, L1Regularizer, ReweightedL1Regularizer, TVRegularizer
from .array_ops import ObjectFunction, Gradient, Mask
from .differentiator import Differentiator
from .graph import GraphedDifferentiator
from .ptychography import reconstruct_ptychography

__all__ = ['multislice_propagate_batch', 'get_kernel', 'gen_freq_mesh', 'apply_rotation', 'pad_object',
           'calculate_pad_len', 'save_rotation_lookup', 'read_origin_coords', 'initialize_probe',
           'ForwardModel', 'PtychographyModel', 'SingleBatchFullfieldModel', 'SingleBatchPtychographyModel',
           'SparseMultisliceModel', 'MultiDistModel',
           'Optimizer', 'AdamOptimizer', 'GDOptimizer', 'MomentumOptimizer', 'Regularizer', 'L1Regularizer', 'ReweightedL1Regularizer',
           'TVRegularizer', 'ObjectFunction', 'Gradient', 'Mask', 'Differentiator', 'GraphedDifferentiator', 'reconstruct_ptychography']
