"""GPU parity of the mixture-of-experts variants: `ggml_moe_a8_vec` (grouped MMVQ, one
launch for every (token, expert) pair) and `ggml_moe_a8` (per-expert MMQ over `moe_align_block_size` token lists).

The reference has no MoE op (the ops are its upstream's, vLLM `ggml_moe_a8_vec` / `ggml_moe_a8`), so the oracle is the
definition of the op over the pinned single-matrix oracles: pair i = token i // top_k through expert ids[i], i.e.
`oracle.mmvq(W[ids[i]], x[i // top_k])` resp. `oracle.mmq(...)`, at the tolerances of tests/test_gpu_mmvq.py and
tests/test_gpu_mmq.py.  Where the plain op runs the same kernel body at batch 1 (Q5_1 and the codebook IQ types) the
grouped result must equal `ggml_mul_mat_vec_a8(W[e], x[t])` bit for bit."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

TYPES = oracle.VEC_TYPES
IDS = [oracle.TYPE_NAMES[t] for t in TYPES]
OUT_REL = {torch.float32: 0.0, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}
# batch 1 of these types is served by the generic kernel (csrc/mmvq.cu) in ggb_mmvq too: same body, same summation order
GENERIC_B1_TYPES = [t for t in TYPES if t in (7, 16, 17, 18, 19, 21, 22, 29)]


@pytest.fixture(scope="module")
def ops():
    import ggml
    return ggml


def _experts(t, E, N, K, seed):
    return np.stack([oracle.random_weights(t, N, K, seed=seed + 31 * e + t) for e in range(E)])


def _vec_case(ops, t, E, N, K, tokens, top_k, dtype, seed=0, ids=None):
    w = _experts(t, E, N, K, seed)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(tokens, K, generator=g).to(dtype)
    if ids is None:
        ids = torch.randint(0, E, (tokens, top_k), generator=g, dtype=torch.int64)
    got = ops.ggml_moe_a8_vec(x.cuda(), torch.from_numpy(w).cuda(), ids.cuda(), top_k, t, N, tokens)
    assert got.shape == (tokens * top_k, N) and got.dtype == dtype
    got = got.float().cpu().numpy()
    flat = ids.reshape(-1).tolist()
    x32 = x.float().numpy()
    for i, e in enumerate(flat):
        tok = i // top_k
        want = oracle.mmvq(w[e], t, x32[tok:tok + 1], N)[0]
        scale = float(np.abs(want).max()) + 1e-30
        tol = 2e-5 * scale + OUT_REL[dtype] * np.abs(want) * 1.01
        bad = np.abs(got[i] - want) > tol
        assert not bad.any(), (f"{oracle.TYPE_NAMES[t]} pair {i} (token {tok}, expert {e}) N={N} K={K} {dtype}: "
                               f"{bad.sum()} of {bad.size} off; max err {np.abs(got[i] - want).max():.3e} scale {scale:.3e}")
    return w, x, ids, got


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_moe_vec_vs_oracle(ops, t, dtype):
    _vec_case(ops, t, E=4, N=97, K=2304, tokens=3, top_k=2, dtype=dtype, seed=1)


@pytest.mark.parametrize("t", TYPES, ids=IDS)
def test_moe_vec_shapes(ops, t):
    _vec_case(ops, t, E=3, N=1, K=256, tokens=1, top_k=1, dtype=torch.float16, seed=2)        # smallest
    _vec_case(ops, t, E=8, N=300, K=4096, tokens=2, top_k=4, dtype=torch.float16, seed=3)     # two full chunks
    _vec_case(ops, t, E=2, N=40, K=512, tokens=90, top_k=2, dtype=torch.float32, seed=4)      # more pairs than SMs


@pytest.mark.parametrize("t", oracle.BLOCK32_TYPES, ids=[oracle.TYPE_NAMES[t] for t in oracle.BLOCK32_TYPES])
def test_moe_vec_unaligned_rows(ops, t):
    """legacy types with K % 32 == 0 only: rows (and expert matrices) are not 16-byte aligned -> the plain-copy path."""
    _vec_case(ops, t, E=3, N=9, K=96, tokens=2, top_k=2, dtype=torch.float32, seed=5)
    _vec_case(ops, t, E=2, N=17, K=2080, tokens=1, top_k=2, dtype=torch.float16, seed=6)


@pytest.mark.parametrize("t", GENERIC_B1_TYPES, ids=[oracle.TYPE_NAMES[t] for t in GENERIC_B1_TYPES])
def test_moe_vec_equals_plain_op_bitwise(ops, t):
    E, N, K, tokens, top_k = 4, 130, 2560, 3, 2
    w, x, ids, got = _vec_case(ops, t, E, N, K, tokens, top_k, torch.float32, seed=7)
    wd = torch.from_numpy(w).cuda()
    for i, e in enumerate(ids.reshape(-1).tolist()):
        tok = i // top_k
        plain = ops.ggml_mul_mat_vec_a8(wd[e], x[tok:tok + 1].cuda(), t, N).cpu().numpy()[0]
        assert np.array_equal(plain.view(np.uint32), got[i].view(np.uint32)), f"pair {i} differs from the plain op"


def test_moe_vec_down_projection_form(ops):
    """second projection of an expert MLP: top_k = 1, one activation row per pair."""
    t, E, N, K, pairs = oracle.Q4_K, 4, 64, 1024, 6
    ids = torch.tensor([[3], [0], [0], [2], [1], [3]])
    _vec_case(ops, t, E, N, K, pairs, 1, torch.float16, seed=8, ids=ids)


def test_moe_vec_out_of_range_ids_give_zero_rows(ops):
    t, E, N, K = oracle.Q6_K, 3, 50, 512
    ids = torch.tensor([[0, -1], [3, 2], [1, 7]])              # -1, 3 (== E) and 7 are outside [0, E)
    w = _experts(t, E, N, K, 9)
    x = torch.randn(3, K, generator=torch.Generator().manual_seed(9)).to(torch.float16)
    got = ops.ggml_moe_a8_vec(x.cuda(), torch.from_numpy(w).cuda(), ids.cuda(), 2, t, N, 3).float().cpu().numpy()
    for i, e in enumerate(ids.reshape(-1).tolist()):
        if 0 <= e < E:
            want = oracle.mmvq(w[e], t, x.float().numpy()[i // 2:i // 2 + 1], N)[0]
            assert np.abs(got[i] - want).max() <= 2e-5 * np.abs(want).max() + 2.0 ** -10 * 1.01 * np.abs(want).max()
        else:
            assert not got[i].any(), f"pair {i} (expert id {e}) must be zero"


def test_moe_vec_guard_region_and_int32_ids(ops):
    """the out-variant writes exactly [pairs, N] (canaries either side stay) and takes int32 ids as they are."""
    t, E, N, K, tokens, top_k = oracle.Q4_0, 4, 33, 256, 5, 2
    w = _experts(t, E, N, K, 10)
    x = torch.randn(tokens, K, generator=torch.Generator().manual_seed(10)).to(torch.float32)
    ids = torch.randint(0, E, (tokens * top_k,), generator=torch.Generator().manual_seed(11), dtype=torch.int32)
    pad = 64
    buf = torch.full((pad + tokens * top_k * N + pad,), 12345.0, dtype=torch.float32, device="cuda")
    out = buf[pad:pad + tokens * top_k * N].view(tokens * top_k, N)
    ops.moe_a8_vec_out(x.cuda(), torch.from_numpy(w).cuda(), ids.cuda(), top_k, t, N, tokens, out)
    torch.cuda.synchronize()
    b = buf.cpu().numpy()
    assert (b[:pad] == 12345.0).all() and (b[-pad:] == 12345.0).all(), "wrote outside the output"
    got = out.cpu().numpy()
    for i, e in enumerate(ids.tolist()):
        want = oracle.mmvq(w[e], t, x.numpy()[i // top_k:i // top_k + 1], N)[0]
        assert np.abs(got[i] - want).max() <= 2e-5 * np.abs(want).max()


def test_moe_vec_mixtral_shape(ops):
    """8 experts of the Llama-8B / Mixtral FFN shape (Q4_K [14336, 4096] each, 264 MB), one token, top-2."""
    _vec_case(ops, oracle.Q4_K, E=8, N=14336, K=4096, tokens=1, top_k=2, dtype=torch.float16, seed=12,
              ids=torch.tensor([[5, 2]]))


def test_moe_vec_more_pairs_than_one_grid(ops):
    """80000 pairs > the 65535 limit of gridDim.y: the op splits into launches of whole tokens; the result equals the
    same op applied to two halves of the tokens, bit for bit."""
    t, E, N, K, tokens, top_k = oracle.Q8_0, 3, 2, 256, 40000, 2
    w = torch.from_numpy(_experts(t, E, N, K, 15)).cuda()
    g = torch.Generator().manual_seed(15)
    x = torch.randn(tokens, K, generator=g).to(torch.float16).cuda()
    ids = torch.randint(0, E, (tokens, top_k), generator=g, dtype=torch.int32).cuda()
    whole = ops.ggml_moe_a8_vec(x, w, ids, top_k, t, N, tokens)
    h = 20000
    parts = torch.cat([ops.ggml_moe_a8_vec(x[:h], w, ids[:h], top_k, t, N, h),
                       ops.ggml_moe_a8_vec(x[h:], w, ids[h:], top_k, t, N, tokens - h)])
    assert torch.equal(whole.view(torch.int16), parts.view(torch.int16))
    # and a few pairs either side of the launch boundary against the oracle
    xs, ws, idl = x.float().cpu().numpy(), w.cpu().numpy(), ids.reshape(-1).cpu().tolist()
    for i in (0, 65533, 65534, 65535, 79999):
        want = oracle.mmvq(ws[idl[i]], t, xs[i // top_k:i // top_k + 1], N)[0]
        assert np.abs(whole[i].float().cpu().numpy() - want).max() <= (2e-5 + 2.0 ** -10 * 1.01) * np.abs(want).max()


def test_moe_vec_rejects_bad_arguments(ops):
    t, E, N, K = oracle.Q4_K, 2, 8, 256
    w = torch.from_numpy(_experts(t, E, N, K, 13)).cuda()
    x = torch.zeros(2, K, dtype=torch.float16, device="cuda")
    ids = torch.zeros(2, 2, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        ops.ggml_moe_a8_vec(x, w[0], ids, 2, t, N, 2)                      # W not an expert stack
    with pytest.raises(ValueError):
        ops.ggml_moe_a8_vec(x, w, ids[:1], 2, t, N, 2)                     # too few ids
    with pytest.raises(ValueError):
        ops.ggml_moe_a8_vec(x, w, ids, 2, t, N + 1, 2)                     # row does not match W
    with pytest.raises(ValueError):
        ops.ggml_moe_a8_vec(x.cpu(), w, ids, 2, t, N, 2)                   # no CPU path


# ------------------------------------------------------------------------------------------------ prefill: ggml_moe_a8

MMQ_CASES = [(oracle.Q4_K, torch.float16), (oracle.Q6_K, torch.bfloat16), (oracle.Q8_0, torch.float16),
             (oracle.Q4_0, torch.float32), (oracle.Q2_K, torch.float16)]


@pytest.mark.parametrize("t,dtype", MMQ_CASES, ids=[oracle.TYPE_NAMES[t] for t, _ in MMQ_CASES])
def test_moe_a8_vs_oracle(ops, t, dtype):
    E, N, K, tokens, top_k = 5, 136, 512, 70, 2
    w = _experts(t, E, N, K, 14)
    g = torch.Generator().manual_seed(14)
    x = (torch.randn(tokens, K, generator=g) * 0.25).to(dtype)
    ids = torch.randint(0, E - 1, (tokens, top_k), generator=g, dtype=torch.int64)       # expert E-1 gets no token
    block = ops.ggml_moe_get_block_size(t)
    sorted_ids, expert_ids, n_pad = ops.moe_align_block_size(ids.cuda(), block, E)
    got = ops.ggml_moe_a8(x.cuda(), torch.from_numpy(w).cuda(), sorted_ids, expert_ids, n_pad, t, N, top_k, tokens)
    assert got.shape == (tokens * top_k, N) and got.dtype == dtype
    got = got.float().cpu().numpy()
    x32 = x.float().numpy()
    flat = ids.reshape(-1).numpy()
    out_rel = {torch.float32: 0.0, torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}[dtype]
    for e in range(E):
        pairs = np.nonzero(flat == e)[0]
        if len(pairs) == 0:
            continue
        want = oracle.mmq(w[e], t, x32[pairs // top_k], N)
        # tests/test_gpu_mmq.py: fp16 operands on the tensor cores, 2^-9 of the scale + one output rounding
        tol = 2.0 ** -9 * np.abs(want).max() + 1.01 * out_rel * np.abs(want)
        assert (np.abs(got[pairs] - want) <= tol).all(), f"expert {e}: max err {np.abs(got[pairs] - want).max():.3e}"


def test_moe_a8_rejects_types_without_mmq(ops):
    with pytest.raises(ValueError):
        ops.ggml_moe_get_block_size(oracle.IQ4_XS if hasattr(oracle, "IQ4_XS") else 23)
