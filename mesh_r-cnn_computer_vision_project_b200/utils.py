"""Mirror of the hot-path helpers of reference ``meshRCNN/utils.py``."""
import torch
from torch import Tensor

from . import functional as F_


def aggregate_neighbours(index: Tensor, matrix: Tensor) -> Tensor:
    """out[row] += matrix[col] over the 2 x E COO edge list ``index`` -- reference meshRCNN/utils.py:52-57
    (there: gather + ``scatter_add_`` atomics; here: deterministic CSR row gather)."""
    return F_.aggregate_neighbours(index, matrix)


def dummy(*dims) -> Tensor:
    """Deterministic test tensor (reference meshRCNN/utils.py:103-109)."""
    n = 1
    for d in dims:
        n *= d
    return torch.arange(n).float().reshape(*dims)
