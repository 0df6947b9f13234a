"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU pool, profiles/r02_compute_sanitizer.txt):
every op writes its output -- and its workspace -- inside a larger allocation whose surroundings are filled with a sentinel;
after the call the guard regions must be untouched and the result must equal the one computed in a plain allocation."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
GUARD = 4096          # bytes on each side
SENT = 0xA5


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _guarded(nbytes, dtype, shape):
    """a tensor of `shape` / `dtype` viewed into the middle of a sentinel-filled byte buffer (256-byte aligned start)"""
    raw = torch.full((GUARD + nbytes + GUARD,), SENT, dtype=torch.uint8, device="cuda")
    return raw, raw[GUARD:GUARD + nbytes].view(dtype).view(*shape)


def _intact(raw, nbytes):
    return bool((raw[:GUARD] == SENT).all()) and bool((raw[GUARD + nbytes:] == SENT).all())


@pytest.mark.parametrize("t", oracle.VEC_TYPES, ids=[oracle.TYPE_NAMES[t] for t in oracle.VEC_TYPES])
def test_dequantize_stays_inside(ops, t):
    for m, n in ((33, 1024), (7, 256)):
        w = torch.from_numpy(oracle.random_weights(t, m, n, seed=1)).cuda()
        for dt in (torch.float16, torch.float32):
            nb = m * n * dt.itemsize
            raw, out = _guarded(nb, dt, (m, n))
            ops.dequantize_out(w, t, m, n, out)
            torch.cuda.synchronize()
            assert _intact(raw, nb), f"dequantize {oracle.TYPE_NAMES[t]} {m}x{n} {dt} wrote outside its output"
            assert torch.equal(out, ops.ggml_dequantize(w, t, m, n, dtype=dt))


@pytest.mark.parametrize("B", [1, 2, 3, 8])
@pytest.mark.parametrize("case", [(oracle.Q4_K, 301, 4096), (oracle.Q4_K, 67, 14336), (oracle.Q6_K, 129, 2048), (oracle.Q5_K, 65, 1024),
                                  (oracle.IQ4_XS, 66, 1024), (oracle.Q8_0, 41, 288), (oracle.Q2_K, 70, 512), (oracle.Q4_0, 33, 1056)],
                         ids=lambda c: f"{oracle.TYPE_NAMES[c[0]]}-{c[1]}x{c[2]}")
def test_mmvq_stays_inside(ops, case, B):
    t, N, K = case
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=2)).cuda()
    x = torch.randn(B, K, device="cuda", dtype=torch.float16)
    nb = B * N * 2
    raw, out = _guarded(nb, torch.float16, (B, N))
    need = max(ops.mmvq_workspace_bytes(t, B, K, N), 16)
    need = (need + 255) // 256 * 256
    wraw, ws = _guarded(need, torch.uint8, (need,))
    ops.mul_mat_vec_a8_out(w, x, t, N, out, ws)
    torch.cuda.synchronize()
    assert _intact(raw, nb), "MMVQ wrote outside its output"
    assert _intact(wraw, need), "MMVQ wrote outside its workspace"
    assert torch.equal(out, ops.ggml_mul_mat_vec_a8(w, x, t, N))


@pytest.mark.parametrize("case", [(oracle.Q4_K, 300, 1024, 83), (oracle.Q6_K, 129, 2048, 7), (oracle.Q8_0, 130, 2080, 20), (oracle.Q4_K, 19008, 512, 520),
                                  (oracle.Q2_K, 128, 256, 300)], ids=lambda c: f"{oracle.TYPE_NAMES[c[0]]}-{c[1]}x{c[2]}-M{c[3]}")
def test_mmq_stays_inside(ops, case):
    t, N, K, M = case
    w = torch.from_numpy(oracle.random_weights(t, N, K, seed=3)).cuda()
    x = torch.randn(M, K, device="cuda", dtype=torch.float16)
    nb = M * N * 2
    raw, out = _guarded(nb, torch.float16, (M, N))
    need = (max(ops.mmq_workspace_bytes(t, M, K, N), 256) + 255) // 256 * 256
    wraw, ws = _guarded(need, torch.uint8, (need,))
    ops.mul_mat_a8_out(w, x, t, N, out, ws)
    torch.cuda.synchronize()
    assert _intact(raw, nb), "MMQ wrote outside its output"
    assert _intact(wraw, need), "MMQ wrote outside its workspace"
    assert torch.equal(out, ops.ggml_mul_mat_a8(w, x, t, N))


def test_q8_1_stays_inside(ops):
    x = torch.randn(9, 1056, device="cuda", dtype=torch.float16)
    for lay in (ops.GGB_Q8_1, ops.GGB_Q8_1_MMQ):
        nb = ops.q8_1_bytes(9, 1056, lay)
        raw, out = _guarded(nb, torch.uint8, (nb,))
        from ggml import _capi
        _capi.check(_capi.lib.ggb_quantize_q8_1(x.data_ptr(), ops.GGB_F16, 9, 1056, out.data_ptr(), lay, 12, torch.cuda.current_stream().cuda_stream), "q8_1")
        torch.cuda.synchronize()
        assert _intact(raw, nb)
        assert torch.equal(out, ops.ggml_quantize_q8_1(x, lay, 12).reshape(-1))
