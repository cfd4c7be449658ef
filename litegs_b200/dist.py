"""Data-parallel plumbing for the render path: views shard across ranks, parameters are replicated, and the
only exchange is one all-reduce of the dense per-Gaussian gradient buffer per step.

The reference has no distributed code at all; this is new functionality built on
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).  Nothing here computes gradients: it only
places the chunk-compacted gradients of each rendered view into a flat ``[rows, chunks, chunk_size]``
buffer (views see different visible-chunk sets, so the sum over views is dense) and reduces it.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence

import torch

PARAM_ORDER = ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Round-robin: rank r renders views r, r+world, ... (every rank gets the same count +-1)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_views, world))


def param_rows(shapes: Dict[str, Sequence[int]]) -> Dict[str, slice]:
    """Row ranges of each parameter inside the flat buffer; a parameter [a,b,...,C,S] takes prod(a,b,...) rows."""
    out, r = {}, 0
    for k in PARAM_ORDER:
        lead = 1
        for d in shapes[k][:-2]:
            lead *= int(d)
        out[k] = slice(r, r + lead)
        r += lead
    return out


class GradAccumulator:
    """Dense gradient buffer [rows, C, S] (fp32): sum of the compacted per-view gradients of this rank, then
    all-reduced (sum) over ranks.  59 rows at sh_degree 3 => 236 B per Gaussian."""

    def __init__(self, params: Dict[str, torch.Tensor], symmetric: Optional[bool] = None):
        """symmetric: map the buffer for the own NVLS all-reduce kernel (a COLLECTIVE allocation: every rank must construct the
        accumulator).  None = env LGS_NVLS=1 (default off: ncclAllReduce); False for a rank-local buffer."""
        self._want_symmetric = (os.environ.get("LGS_NVLS", "0") == "1") if symmetric is None else bool(symmetric)
        self.shapes = {k: tuple(params[k].shape) for k in PARAM_ORDER}
        self.rows = param_rows(self.shapes)
        C, S = self.shapes["xyz"][-2:]
        n_rows = self.rows["opacity"].stop
        dev = params["xyz"].device
        # one allocation = one all-reduce: [rows*C*S gradient | C chunk marks].  The marks say which chunks were visible
        # in at least one view of the step (on any rank after the reduce); the fused optimizer step updates only those,
        # the reference's sparse-Adam semantics (optimizer.py:14-44).
        n_flat = n_rows * C * S + C
        self._nvls = None                   # symmetric-memory handle when the buffer is mapped for NVSwitch multicast
        self.flat_all = self._allocate(n_flat, dev)
        self.flat = self.flat_all[:n_flat]
        self.buf = self.flat[: n_rows * C * S].view(n_rows, C, S)
        self.touched = self.flat[n_rows * C * S:]
        self._work = None
        self._comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def _allocate(self, n_flat: int, dev) -> torch.Tensor:
        """The flat buffer.  On a multi-GPU NCCL job it is symmetric memory mapped for NVSwitch multicast when the platform offers
        it and the caller asks for it (torch.distributed._symmetric_memory; LGS_NVLS=1 or symmetric=True), so that all_reduce() can run the library's own NVLS kernel
        (csrc/nvls.cu) instead of ncclAllReduce; otherwise an ordinary allocation reduced by the backend's all-reduce."""
        import torch.distributed as dist
        n_pad = (n_flat + 3) // 4 * 4
        if (dev.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and dist.get_backend() == "nccl" and self._want_symmetric):
            try:
                import torch.distributed._symmetric_memory as symm
                t = symm.empty(n_pad, dtype=torch.float32, device=dev)
                hdl = symm.rendezvous(t, dist.group.WORLD)
                if int(hdl.multicast_ptr) == 0:
                    raise RuntimeError("no multicast support")
                t.zero_()
                self._nvls = hdl
                return t
            except Exception as e:      # noqa: BLE001 -- any failure of the optional fast path falls back to NCCL, loudly once
                if os.environ.get("RANK", "0") == "0":
                    print(f"[litegs_b200.dist] NVLS all-reduce unavailable ({type(e).__name__}: {str(e)[:120]}); using the NCCL all-reduce")
                self._nvls = None
        return torch.zeros(n_pad, dtype=torch.float32, device=dev)

    def zero_(self):
        self.flat_all.zero_()

    def mark(self, chunk_ids: torch.Tensor, visible_count: torch.Tensor):
        """touched[chunk_ids[j]] = 1 for j < *visible_count (device count: no sync)."""
        if self.touched.is_cuda:
            from . import _lib
            from .fused import _ptr, _stream
            _lib.call("lgs_mark_visible_chunks", _ptr(chunk_ids), _ptr(visible_count), int(chunk_ids.shape[0]), _ptr(self.touched),
                      _stream(self.touched.device))
        else:
            n = int(visible_count.reshape(-1)[0])
            self.touched[chunk_ids[:n].long()] = 1.0

    def add_view(self, compacted: Dict[str, torch.Tensor], chunk_ids: torch.Tensor, visible_count: torch.Tensor):
        """buf[rows(k), chunk_ids[j], :] += compacted[k][..., j, :] for j < *visible_count (device count: no sync)."""
        self.mark(chunk_ids, visible_count)
        for k in PARAM_ORDER:
            g = compacted[k]
            if g.numel() == 0:
                continue
            A, S = g.shape[-2:]
            g3 = g.reshape(-1, A, S)
            dst = self.buf[self.rows[k]]
            if dst.is_cuda:
                from . import _lib
                from .fused import _ptr, _stream
                if not g3.is_contiguous():
                    g3 = g3.contiguous()
                _lib.call("lgs_sparse_chunk_op", ctypes.c_void_p(dst.data_ptr()), _ptr(g3), _ptr(chunk_ids), _ptr(visible_count), 0, 0,
                          g3.shape[0], dst.shape[1], A, S, _stream(dst.device))
            else:   # host-side logic under gloo (tests): same semantics with torch indexing
                n = int(visible_count.reshape(-1)[0])
                dst[:, chunk_ids[:n].long(), :] += g3[:, :n, :]

    def all_reduce(self, async_op: bool = False):
        """Sum over ranks.  With async_op the collective runs on a side stream so that the next micro-batch's
        forward overlaps it; call wait() before reading the buffer."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if self._nvls is not None:
            if self._comm_stream is not None and async_op:
                self._comm_stream.wait_stream(torch.cuda.current_stream(self.buf.device))
                with torch.cuda.stream(self._comm_stream):
                    self._nvls_all_reduce()
            else:
                self._nvls_all_reduce()
            return
        if self._comm_stream is not None and async_op:
            self._comm_stream.wait_stream(torch.cuda.current_stream(self.buf.device))
            with torch.cuda.stream(self._comm_stream):
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)

    def _nvls_all_reduce(self):
        """barrier (every replica complete) -> one multimem kernel: each rank reduces its slice inside the switch and broadcasts it
        -> barrier (every slice delivered), all on the current stream."""
        from . import _lib
        hdl = self._nvls
        dev = self.buf.device
        st = torch.cuda.current_stream(dev).cuda_stream
        hdl.barrier(channel=0, timeout_ms=60000)
        _lib.call("lgs_nvls_allreduce_f32", ctypes.c_void_p(int(hdl.multicast_ptr)), ctypes.c_size_t(self.flat_all.numel()), int(hdl.rank),
                  int(hdl.world_size), int(os.environ.get("LGS_NVLS_CTAS", "0")), st)
        hdl.barrier(channel=0, timeout_ms=60000)

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.buf.device).wait_stream(self._comm_stream)

    def grads(self) -> Dict[str, torch.Tensor]:
        """Views of the buffer with the parameters' own shapes, plus "_touched" (f32[C] chunk marks): the dict
        render_view / render_views take as accumulate_into."""
        out = {k: self.buf[self.rows[k]].reshape(self.shapes[k]) for k in PARAM_ORDER}
        out["_touched"] = self.touched
        return out
