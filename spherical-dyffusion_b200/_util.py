"""Small host-side helpers shared by the modules: stream handles, workspace cache, pointer access."""
from __future__ import annotations

import torch

_workspaces: dict = {}


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(device, nbytes: int, tag: str = "default") -> torch.Tensor:
    """Per-(device, tag) scratch buffer, grown on demand; 1024-byte aligned (torch's allocator gives 512+,
    so over-allocate and slice)."""
    key = (str(device), tag)
    buf = _workspaces.get(key)
    need = int(nbytes) + 1024
    if buf is None or buf.numel() < need:
        buf = None
        _workspaces.pop(key, None)
        buf = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + int(nbytes)]


def release_workspaces() -> None:
    _workspaces.clear()


def require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback (got device {t.device})")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
