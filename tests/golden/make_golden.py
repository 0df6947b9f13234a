"""Generate tests/golden/dequant_golden.npz from the REFERENCE implementations, in the build container.

Sources of truth (nothing from oracle/ggml_oracle.c is used here):
  * gguf-py `gguf.quants.dequantize` (gguf 0.19.0) -- the expected-value source of every reference
    test (tests/test_dequantize.py:66, hf-kernels/ggml-kernels/tests/kernels/test_cuda_kernels.py:51-53);
  * the reference's own CPU op `custom_ops.ggml_dequantize` compiled unmodified from
    /root/reference/ggml-cpu/custom_ops.cpp (oracle/build_ref.py) for the five legacy types.

Weights: `gguf.quants.quantize(randn * 0.02)` for the legacy types (the reference fixtures are made
that way) and random valid block bytes for the K-quants and the IQ types (gguf-py cannot quantize them), plus
hand-built edge-case blocks.  IQ3_S: gguf-py's codebook is first moved to the reference's (top level 15.5 instead of 15;
tests/util.py patch_gguf_iq3s_to_reference_table).  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from gguf import GGMLQuantizationType as T, quants  # noqa: E402

import oracle  # noqa: E402  (only for random_weights / block geometry / building the reference op)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import patch_gguf_iq3s_to_reference_table  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dequant_golden.npz")


def edge_blocks(t):
    """A few adversarial blocks: zero / max payloads, d = 0, negative, subnormal and max-normal fp16 scales."""
    qk, bs = oracle.block_elems(t), oracle.block_bytes(t)
    offs = {2: [0], 3: [0, 2], 6: [0], 7: [0, 2], 8: [0], 10: [80, 82], 11: [108], 12: [0, 2], 13: [0, 2], 14: [208],
            20: [0], 23: [0], 16: [0], 17: [0], 18: [0], 19: [0], 21: [0], 22: [0], 29: []}[t]
    halves = np.array([0.0, -0.0, 1.0, -1.0, 6.0e-8, -6.0e-8, 65504.0, 0.333251953125, 1e-3, -2.5], dtype=np.float16)
    rng = np.random.default_rng(1234 + t)
    blocks = []
    for fill in (0x00, 0xFF, 0x55, 0xAA, None):
        for h in halves:
            b = np.full(bs, fill, dtype=np.uint8) if fill is not None else rng.integers(0, 256, bs, dtype=np.uint8)
            for o in offs:
                b[o:o + 2] = np.array([h], dtype=np.float16).view(np.uint8)
            if t == 29:          # IQ1_M: the fp16 scale is spread over the top nibbles of bytes 49, 51, 53, 55
                hb = int(np.array([h], dtype=np.float16).view(np.uint16)[0])
                for i in range(4):
                    b[49 + 2 * i] = (int(b[49 + 2 * i]) & 0x0f) | (((hb >> (4 * i)) & 0x0f) << 4)
            blocks.append(b)
    # second scale field different from the first
    if len(offs) == 2:
        for h0, h1 in ((1.0, -0.5), (0.0, 3.0), (2.0 ** -14, 2.0 ** -15)):
            b = rng.integers(0, 256, bs, dtype=np.uint8)
            b[offs[0]:offs[0] + 2] = np.array([h0], dtype=np.float16).view(np.uint8)
            b[offs[1]:offs[1] + 2] = np.array([h1], dtype=np.float16).view(np.uint8)
            blocks.append(b)
    n = len(blocks)
    return np.stack(blocks).reshape(1, n * bs), n * qk


def main():
    patch_gguf_iq3s_to_reference_table()
    ref = oracle.ref_custom_ops()
    assert ref is not None, "the reference CPU op must be built (oracle/build_ref.py) to make golden vectors"
    data = {}
    rng = np.random.default_rng(0)
    for t in oracle.VEC_TYPES:
        name = oracle.TYPE_NAMES[t]
        rows, cols = 6, 512
        if t in oracle.LEGACY_TYPES:
            w = quants.quantize((rng.standard_normal((rows, cols)) * 0.02).astype(np.float32), T(t))
        else:
            w = oracle.random_weights(t, rows, cols, seed=100 + t)
        y = quants.dequantize(w, T(t)).astype(np.float32)
        data[f"{name}_w"] = w
        data[f"{name}_y"] = y
        ew, en = edge_blocks(t)
        ey = quants.dequantize(ew, T(t)).astype(np.float32)
        data[f"{name}_edge_w"] = ew
        data[f"{name}_edge_y"] = ey
        if t in oracle.LEGACY_TYPES:
            r = ref.ggml_dequantize(torch.from_numpy(w.copy()), t, rows, cols).numpy()
            er = ref.ggml_dequantize(torch.from_numpy(ew.copy()), t, 1, en).numpy()
            assert np.array_equal(r.view(np.uint32), y.view(np.uint32)), name
            assert np.array_equal(er.view(np.uint32), ey.view(np.uint32)), name + " edge"
            data[f"{name}_refcpu_y"] = r
    np.savez_compressed(OUT, **data)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
