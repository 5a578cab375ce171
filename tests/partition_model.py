"""Image-space partition used for multi-GPU rendering (host-side index logic, numpy only).

The frame is cut into bands of `block_rows` scanlines; band b belongs to rank b % n_parts. Every rank
renders its bands into a PACKED buffer (band after band) of the same size on every rank, so that a
single NCCL gather (equal counts) brings them to rank 0, where rt_unpack_rows scatters them back.
This file mirrors rt_rows_packed_pixels / k_unpack_rows of librtcore for the CPU-side tests and the
bench harness; the product's device path is the CUDA kernel.
"""
from __future__ import annotations

import numpy as np


def bands_total(height: int, block_rows: int) -> int:
    return (height + block_rows - 1) // block_rows


def packed_rows(height: int, block_rows: int, n_parts: int) -> int:
    return ((bands_total(height, block_rows) + n_parts - 1) // n_parts) * block_rows


def packed_pixels(width: int, height: int, block_rows: int, n_parts: int) -> int:
    return packed_rows(height, block_rows, n_parts) * width


def global_row(local_row: int, block_rows: int, part_index: int, n_parts: int) -> int:
    band = local_row // block_rows
    return (band * n_parts + part_index) * block_rows + local_row % block_rows


def local_rows_of_part(height: int, block_rows: int, part_index: int, n_parts: int) -> np.ndarray:
    """Global row (or -1 for padding rows past the image) of every packed row of this part."""
    lr = np.arange(packed_rows(height, block_rows, n_parts))
    y = ((lr // block_rows) * n_parts + part_index) * block_rows + lr % block_rows
    return np.where(y < height, y, -1)


def unpack(packed_all: np.ndarray, width: int, height: int, block_rows: int, n_parts: int) -> np.ndarray:
    """packed_all: [n_parts, packed_rows, width, C] -> [height, width, C] (numpy mirror of k_unpack_rows)."""
    out = np.zeros((height, width) + packed_all.shape[3:], dtype=packed_all.dtype)
    for p in range(n_parts):
        rows = local_rows_of_part(height, block_rows, p, n_parts)
        ok = rows >= 0
        out[rows[ok]] = packed_all[p, ok]
    return out
