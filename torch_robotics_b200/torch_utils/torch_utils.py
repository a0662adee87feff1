"""Tensor helpers with the reference's names (torch_utils/torch_utils.py:10-53,156-159)."""
import numpy as np
import torch


def get_torch_device(device="cuda"):
    if "cuda" in str(device) and torch.cuda.is_available():
        return torch.device("cuda")
    if "mps" in str(device):
        return torch.device("mps")
    return torch.device("cpu")


DEFAULT_TENSOR_ARGS = {"device": get_torch_device("cuda"), "dtype": torch.float32}


def to_numpy(x, dtype=np.float32, clone=False):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy().astype(dtype)
    if isinstance(x, np.ndarray):
        return x.astype(dtype, copy=clone)
    return np.array(x).astype(dtype)


def to_torch(x, device="cpu", dtype=torch.float, requires_grad=False, clone=False):
    if torch.is_tensor(x):
        return (x.clone() if clone else x).to(device=device, dtype=dtype)
    return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)


def to_torch_2d_min(variable):
    t = to_torch(variable) if isinstance(variable, (np.ndarray, list)) else variable
    return t.unsqueeze(0) if t.ndim == 1 else t


def torch_intersect_1d(a, b):
    """Sorted values present in both 1-D index tensors (each without duplicates)."""
    both, counts = torch.cat([a, b]).unique(return_counts=True)
    return both[counts > 1]
