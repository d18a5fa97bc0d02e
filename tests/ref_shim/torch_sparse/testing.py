"""What the reference's test files import from `torch_sparse.testing` (torch_sparse/testing.py:1-21), for the
reference-suite run against this package: the parameter lists, with the device list reduced to the GPU (this package has
no CPU path), and the small tensor factory."""
import torch

devices = [torch.device("cuda", 0)]
reductions = "sum add mean min max".split()
grad_dtypes = [getattr(torch, name) for name in ("half", "float", "double", "bfloat16")]
dtypes = grad_dtypes[:3] + [torch.int, torch.long] + grad_dtypes[3:]


def tensor(x, dtype, device):
    """`x` as a tensor of `dtype` on `device`; None stays None."""
    if x is None:
        return None
    return torch.tensor(x, dtype=dtype, device=device)
