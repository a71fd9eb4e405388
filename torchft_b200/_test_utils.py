"""Helpers shared by the test-suite (reference: ``torchft/_test_utils.py``).

``gen_splits`` enumerates, for a flat tensor cut into equal chunks, every way of viewing each
chunk as a 2-D matrix — the quantization tests use it to sweep row/column shapes, since the
fp8 wire format scales per 512-element group and must not care how the caller shaped the data.
"""

from __future__ import annotations

import itertools
from typing import Iterable, List, Sequence, Tuple

import torch

Shape = Tuple[int, int]


def any_nan(tensors: Iterable[torch.Tensor]) -> bool:
    """True if any tensor holds a NaN."""
    return any(bool(torch.isnan(t).any()) for t in tensors)


def gen_views(t: torch.Tensor) -> List[Shape]:
    """Every ``(m, n)`` with ``m * n == t.numel()`` and ``m < numel`` (``(1, n)`` only for even sizes,
    matching the reference's enumeration so test matrices line up)."""

    size = t.numel()
    first = 1 if size % 2 == 0 else 2
    return [(m, size // m) for m in range(first, size) if size % m == 0]


def gen_splits(t: torch.Tensor, split_size: int) -> List[List[Shape]]:
    """Cartesian product of :func:`gen_views` over the ``split_size`` chunks of ``t``."""

    per_chunk: Sequence[List[Shape]] = [gen_views(c) for c in torch.split(t, split_size)]
    return [list(combo) for combo in itertools.product(*per_chunk)]
