"""Synthetic GGUF-encoded weights (no checkpoint files are needed or available offline).

`random_blocks` produces random *valid* block bytes for any supported type: every payload byte
uniform in 0..255, every fp16 scale field = fp16(U(0.5, 1.5) * 2^-10) so dequantized values stay
finite and O(1).  Host-side numpy only; nothing here computes the
hot path.
"""
from __future__ import annotations

import numpy as np

BLOCK = {2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34),
         10: (256, 84), 11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210),
         20: (32, 18), 23: (256, 136),
         16: (256, 66), 17: (256, 74), 18: (256, 98), 19: (256, 50), 21: (256, 110), 22: (256, 82), 29: (256, 56)}
HALF_OFFSETS = {2: [0], 3: [0, 2], 6: [0], 7: [0, 2], 8: [0],
                10: [80, 82], 11: [108], 12: [0, 2], 13: [0, 2], 14: [208], 20: [0], 23: [0],
                16: [0], 17: [0], 18: [0], 19: [0], 21: [0], 22: [0], 29: []}
NAMES = {2: "Q4_0", 3: "Q4_1", 6: "Q5_0", 7: "Q5_1", 8: "Q8_0",
         10: "Q2_K", 11: "Q3_K", 12: "Q4_K", 13: "Q5_K", 14: "Q6_K", 20: "IQ4_NL", 23: "IQ4_XS",
         16: "IQ2_XXS", 17: "IQ2_XS", 18: "IQ3_XXS", 19: "IQ1_S", 21: "IQ3_S", 22: "IQ2_S", 29: "IQ1_M"}


def weight_bytes(qtype: int, rows: int, cols: int) -> int:
    qk, bs = BLOCK[int(qtype)]
    return rows * (cols // qk) * bs


def random_blocks(qtype: int, rows: int, cols: int, seed: int = 0) -> np.ndarray:
    """uint8 [rows, cols/QK*block_bytes] of random valid blocks."""
    qtype = int(qtype)
    qk, bs = BLOCK[qtype]
    nb = rows * cols // qk
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, size=(nb, bs), dtype=np.uint8)
    for off in HALF_OFFSETS[qtype]:
        h = (rng.uniform(0.5, 1.5, size=nb) * 2.0 ** -10).astype(np.float16)
        blocks[:, off:off + 2] = h.view(np.uint8).reshape(nb, 2)
    if qtype == 29:         # IQ1_M: the fp16 scale lives in the top nibbles of the four 16-bit scale words (bytes 49, 51, 53, 55)
        h = (rng.uniform(0.5, 1.5, size=nb) * 2.0 ** -10).astype(np.float16).view(np.uint16)
        for i in range(4):
            blocks[:, 49 + 2 * i] = (blocks[:, 49 + 2 * i] & 0x0f) | (((h >> (4 * i)) & 0x0f) << 4).astype(np.uint8)
    return blocks.reshape(rows, cols // qk * bs)
