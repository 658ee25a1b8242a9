"""Small host-side helpers shared by the modules: stream handles, workspace cache, pointer access.

Workspaces are keyed by ``(device, tag)``.  Every native net handle uses its own tag (``net<handle>``), so two modules
(forecaster and interpolator, possibly on different streams) never share scratch memory; the stand-alone ops share one
buffer per op family and therefore assume ONE stream per device at a time.  A buffer that was handed out while a CUDA
graph was being captured is never freed when the workspace grows (the graph has its address baked in): it is parked in
``_retired`` until ``release_workspaces``.
"""
from __future__ import annotations

import torch

_workspaces: dict = {}
_captured: set = set()     # keys whose current buffer was handed out during a stream capture
_retired: list = []        # buffers a captured graph may still reference


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Per-(device, tag) scratch buffer, grown on demand; 1024-byte aligned (torch's allocator gives 512+,
    so over-allocate and slice)."""
    key = (str(device), tag)
    buf = _workspaces.get(key)
    need = int(nbytes) + 1024
    capturing = torch.cuda.is_current_stream_capturing()
    if buf is None or buf.numel() < need:
        if capturing and buf is not None:
            raise RuntimeError(f"workspace {tag!r} must grow from {buf.numel()} to {need} bytes during a CUDA-graph capture: "
                               "run one eager call with the largest batch before capturing")
        if buf is not None and key in _captured:
            _retired.append(buf)   # a captured graph replays into this address: keep it alive
            _captured.discard(key)
        _workspaces.pop(key, None)
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    if capturing:
        _captured.add(key)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + int(nbytes)]


def release_workspace(device, tag: str) -> None:
    key = (str(device), tag)
    buf = _workspaces.pop(key, None)
    if buf is not None and key in _captured:
        _retired.append(buf)
        _captured.discard(key)


def release_workspaces() -> None:
    """Drop every scratch buffer (also the ones parked for captured graphs: destroy those graphs first)."""
    _workspaces.clear()
    _captured.clear()
    _retired.clear()


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback (got device {t.device})")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
