"""tenet.jl_b200 — B200-native drop-in for the contraction hot path of bsc-quantic/Tenet.jl.

The directory name carries a dot, so it is imported through the root-level shim `tenet_jl_b200.py`
(`import tenet_jl_b200 as tb`).  Public names mirror what Tenet.jl re-exports from Muscle/Tangles
(/root/reference/src/Tenet.jl:9-12,27-81) for this path: Tensor, Index, binary_einsum, TensorNetwork, einexpr,
contract, MPS, MPO, ProductState, PEPS, overlap, ising_1d_mpo.
"""
from ._lib import LIB_PATH, TnbError, load_library  # noqa: F401
from .context import B200Array, Context, default_context  # noqa: F401
from .tensor import Index, Tensor, binary_einsum, tensor_qr_thin, tensor_svd_thin  # noqa: F401
from .pathfinder import ContractionPath, find_slices, optimize_path  # noqa: F401
from .network import ContractionPlan, TensorNetwork, contract, einexpr, multi_contract  # noqa: F401
from .components import MPO, MPS, PEPS, ProductState, expect_network, ising_1d_mpo, overlap  # noqa: F401
from . import workloads  # noqa: F401
from . import distributed  # noqa: F401

__all__ = ["Tensor", "Index", "binary_einsum", "tensor_qr_thin", "tensor_svd_thin", "TensorNetwork", "einexpr", "contract", "multi_contract", "ContractionPlan",
           "ContractionPath", "MPS", "MPO", "PEPS", "ProductState", "overlap", "ising_1d_mpo", "expect_network",
           "B200Array", "Context", "default_context", "TnbError", "load_library", "workloads", "distributed"]
