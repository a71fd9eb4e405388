"""Loader for the sm_100a data-plane extension ``torchft_b200._K``.

The extension is built in-tree by ``torchft_b200._build`` (see
``__graft_entry__.build``). On a box with a GPU a missing extension is a hard
error -- there is deliberately no eager-PyTorch fallback for the hot ops, so a
silent slow path can never masquerade as the native one.
"""

from __future__ import annotations

import threading
from typing import Any, Optional

import torch

_lock = threading.Lock()
_K: Optional[Any] = None
_err: Optional[BaseException] = None

DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
OP_SUM, OP_MAX, OP_MIN = 0, 1, 2


def load() -> Any:
    """Return the ``_K`` module, building it on first use if the source tree is newer."""
    global _K, _err
    if _K is not None:
        return _K
    with _lock:
        if _K is not None:
            return _K
        try:
            from torchft_b200 import _K as mod  # type: ignore[attr-defined]
        except ImportError as first:
            try:
                from torchft_b200 import _build

                _build.build_kernels()
                from torchft_b200 import _K as mod  # type: ignore[attr-defined,no-redef]
            except Exception as e:  # pragma: no cover - build toolchain missing
                _err = e
                raise ImportError(
                    "torchft_b200._K (sm_100a kernels) is not built and could not be built: "
                    f"{first!r} / {e!r}. Run `python -m torchft_b200._build`."
                ) from e
        _K = mod
        return mod


def available() -> bool:
    """True when the extension imports AND a CUDA device is present."""
    if not torch.cuda.is_available():
        return False
    try:
        load()
        return True
    except ImportError:
        return False


def stream_ptr(stream: Optional[torch.cuda.Stream] = None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def dtype_code(t: torch.Tensor) -> int:
    try:
        return DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype for native kernels: {t.dtype}") from None
