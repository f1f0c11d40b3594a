"""B200-native engine for TensorNetworkEvolve.jl's hot path (PEPS norm / energy contraction,
their TensorAD gradients, TEBD bond update) -- host-side mirror of the reference's interface over
the C ABI in ``include/tne.h``.  Import as ``tne_b200`` (see ``tne_b200.py`` at the repo root; the
directory name carries the reference's ``.jl`` and is not a Python identifier).

There is no CPU path: every tensor operation is a CUDA kernel in ``libtne_b200.so``; without the
built library or without a B200 the first device call raises.
"""
from . import _lib
from ._lib import TneError, build
from .array import B200Array, Context, context, svd_trunc
from . import einsum, graphs, peps, stateio, tebd, tensorad, timeevolve, yao
from .einsum import EinCode, NestedEinsum, SlicedEinsum, optimize_code
from .graphs import SimpleGraph, grid, graph_from_edges
from .peps import (PEPS, SimplePEPS, VidalPEPS, zero_simplepeps, zero_vidalpeps, rand_simplepeps, rand_vidalpeps,
                   statetensor, statevec, vec, getvlabel, getphysicallabel, newlabel, findbondtensor, virtualbonds,
                   apply_onbond_, apply_onbond_batch_, apply_sweep_, apply_sweep_batch_, apply_onsite_, apply_onsite, inner_product, norm, normalize_, variables,
                   load_variables_, load_variables, expect, nsite, replace_tensors, conj, sweep_plan, slice_bonds, inner_product_sliced)
from .tebd import rydberg_tebd_, rydberg_tebd_sweep_, rydberg_tebd_sweep_batch_, rydberg_tebd_batch_, rydberg_hamiltonian
from .tensorad import DiffTensor, gradient
from .timeevolve import fvec, iloss2, normalized_overlap, apply_smatrix, sr_step, wrap_difftensor, energy_and_gradient
from .stateio import save_peps, load_peps
