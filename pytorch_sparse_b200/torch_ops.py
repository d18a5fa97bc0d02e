"""Registers the hot-path operators with PyTorch's dispatcher under the reference's own schemas
(csrc/spmm.cpp:305-348, csrc/convert.cpp:46-48):

    torch.ops.tsb200.{spmm_sum, spmm_mean, spmm_min, spmm_max, ind2ptr, ptr2ind, cuda_version}

and, unless TSB200_REGISTER_TORCH_SPARSE=0 or a real torch_sparse build already owns the namespace,
the same operators as `torch.ops.torch_sparse.*`, so code that calls the reference's ops by name
keeps working on top of this package.
"""
from __future__ import annotations

import os

import torch

from . import ops

_SCHEMAS = {
    "spmm_sum": "(Tensor? row, Tensor rowptr, Tensor col, Tensor? value, Tensor? colptr, Tensor? csr2csc, "
                "Tensor mat) -> Tensor",
    "spmm_mean": "(Tensor? row, Tensor rowptr, Tensor col, Tensor? value, Tensor? rowcount, Tensor? colptr, "
                 "Tensor? csr2csc, Tensor mat) -> Tensor",
    "spmm_min": "(Tensor rowptr, Tensor col, Tensor? value, Tensor mat) -> (Tensor, Tensor)",
    "spmm_max": "(Tensor rowptr, Tensor col, Tensor? value, Tensor mat) -> (Tensor, Tensor)",
    "ind2ptr": "(Tensor ind, int M) -> Tensor",
    "ptr2ind": "(Tensor ptr, int E) -> Tensor",
    "cuda_version": "() -> int",
}
_IMPLS = {
    "spmm_sum": ops.spmm_sum, "spmm_mean": ops.spmm_mean, "spmm_min": ops.spmm_min, "spmm_max": ops.spmm_max,
    "ind2ptr": ops.ind2ptr, "ptr2ind": ops.ptr2ind, "cuda_version": ops.cuda_version,
}
_LIBS = []  # keep Library objects alive


def _register(ns: str) -> bool:
    try:
        lib = torch.library.Library(ns, "FRAGMENT")
        for name, schema in _SCHEMAS.items():
            lib.define(name + schema)
            # the Python implementations wrap autograd.Functions, i.e. they are "composite implicit"
            lib.impl(name, _IMPLS[name], "CompositeImplicitAutograd")
    except Exception:  # namespace already owned by a compiled torch_sparse
        return False
    _LIBS.append(lib)
    return True


REGISTERED = {"tsb200": _register("tsb200")}
if os.environ.get("TSB200_REGISTER_TORCH_SPARSE", "1") != "0":
    REGISTERED["torch_sparse"] = _register("torch_sparse")
