"""Sparse (chunk-compacted) gradient container.

Plays the role of the reference's ``litegs/utils/CompactedTensor.py``: autograd insists that a gradient
has the metadata (shape / dtype / device) of its leaf, while the render backward only produces values for
the chunks that survived frustum culling.  The wrapper reports the full ``[..., chunks, chunk_size]``
shape and carries ``chunk_ids`` (i64, which chunk each compacted row belongs to) and
``compacted_values`` (``[..., allocated_chunks, chunk_size]``).  Attribute names match the reference so
its ``SparseGaussianAdam`` (training/optimizer.py:37-38) consumes these objects unchanged.
"""
from __future__ import annotations

import torch


class CompactedTensor(torch.Tensor):
    @staticmethod
    def __new__(cls, full_shape, chunk_ids: torch.Tensor, compacted_values: torch.Tensor):
        return torch.Tensor._make_wrapper_subclass(cls, tuple(full_shape), dtype=compacted_values.dtype,
                                                   device=compacted_values.device, layout=torch.strided, requires_grad=False)

    def __init__(self, full_shape, chunk_ids: torch.Tensor, compacted_values: torch.Tensor):
        self.chunk_ids = chunk_ids
        self.compacted_values = compacted_values

    def __repr__(self):
        return f"CompactedTensor(shape={tuple(self.shape)}, compacted_shape={tuple(self.compacted_values.shape)})"

    def to_dense(self, valid_chunks: int | None = None) -> torch.Tensor:
        """Scatter into a zero tensor of the full shape (rows beyond ``valid_chunks`` are ignored)."""
        n = self.chunk_ids.shape[0] if valid_chunks is None else int(valid_chunks)
        out = torch.zeros(tuple(self.shape), dtype=self.dtype, device=self.compacted_values.device)
        out[..., self.chunk_ids[:n], :] = self.compacted_values[..., :n, :]
        return out

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        packet = getattr(func, "overloadpacket", None)
        src = args[0]
        if packet is torch.ops.aten.detach or packet is torch.ops.aten.alias:
            return cls(src.shape, src.chunk_ids, src.compacted_values)
        if packet is torch.ops.aten.clone:
            return cls(src.shape, src.chunk_ids.clone(), src.compacted_values.clone())
        raise NotImplementedError(f"CompactedTensor supports detach/clone only (got {func}); use .to_dense() or .compacted_values")
