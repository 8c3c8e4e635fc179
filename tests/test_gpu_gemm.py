"""tcgen05 GEMM parity (through the C ABI) against torch fp32 matmul on the same bf16 inputs, and against the
SIMT cross-check kernel.  Own process: a trap in this kernel must not poison the other GPU tests."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def ops():
    from distil_whisper_b200 import ops as o, _abi
    _abi.call("dwb_check_device")
    return o


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(torch.bfloat16)


def _ref(a, b, a_mn, b_mn, bias, act, alpha):
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    y = alpha * (A @ Bm)
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    return y


def _rel(x, y):
    return float((x.float() - y).norm() / (y.norm() + 1e-20))


SHAPES = [(128, 256, 64), (256, 512, 320), (304, 136, 240), (1000, 520, 1280), (4096, 1280, 1280), (72, 1288, 200),
          (4096, 1288, 192)]      # the last two 4096-row shapes select the 128x192 tile (with and without an N tail)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts_bf16_out(ops, M, N, K, a_mn, b_mn):
    a = _mk((K, M) if a_mn else (M, K), 1, 0.5)
    b = _mk((K, N) if b_mn else (N, K), 2, 0.5)
    bias = torch.randn(N, device="cuda")
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, act=0)
    ref = _ref(a, b, a_mn, b_mn, bias, 0, 1.0)
    assert _rel(out, ref) < 6e-3, _rel(out, ref)
    chk = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, act=0, impl=1)
    assert _rel(out, chk.float()) < 6e-3


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (1500, 640, 384), (333, 5120, 1280)])
def test_gemm_gelu_epilogue(ops, M, N, K):
    a, b = _mk((M, K), 3, 0.3), _mk((N, K), 4, 0.3)
    bias = torch.randn(N, device="cuda") * 0.1
    out = ops.gemm(a, b, bias=bias, act=1)
    ref = _ref(a, b, False, False, bias, 1, 1.0)
    assert _rel(out, ref) < 8e-3


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (False, True)])
def test_gemm_f32_out_alpha_accumulate(ops, a_mn, b_mn):
    M, N, K = 384, 264, 4096      # long K, few tiles -> split-K path
    a = _mk((K, M) if a_mn else (M, K), 5, 0.2)
    b = _mk((K, N) if b_mn else (N, K), 6, 0.2)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, alpha=0.5)
    ref = _ref(a, b, a_mn, b_mn, None, 0, 0.5)
    assert out.dtype == torch.float32 and _rel(out, ref) < 2e-5, _rel(out, ref)
    base = torch.randn(M, N, device="cuda")
    acc = base.clone()
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=acc, alpha=0.5, accumulate=True)
    assert _rel(acc, base + ref) < 2e-5


def test_gemm_vocab_tail_f32(ops):
    # LM-head shape class: N not a multiple of the tile, fp32 logits into a padded-pitch buffer
    M, N, K = 96, 6483, 128
    a, b = _mk((M, K), 7), _mk((N, K), 8)
    out = ops.gemm(a, b, out_dtype=torch.float32)
    assert out.shape == (M, N) and out.stride(0) % 4 == 0
    assert _rel(out, _ref(a, b, False, False, None, 0, 1.0)) < 2e-5


def test_gemm_rejects_misaligned_pitch(ops):
    from distil_whisper_b200._abi import DwbError
    a, b = _mk((64, 68), 9), _mk((64, 68), 10)      # pitch 68 * 2 B = 136 B: not a multiple of 16 B
    with pytest.raises(DwbError):
        ops.gemm(a[:, :64], b[:, :64])


PAIR_SHAPES = [(256, 256, 64), (512, 512, 320), (304, 136, 240), (1000, 520, 1280), (4096, 1280, 1280), (72, 1288, 200), (2000, 3840, 384)]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_gemm_cta_pair_layouts(ops, M, N, K, a_mn, b_mn):
    """cta_group::2 kernel forced (impl=2): 256-row MMAs issued by the pair leader, B tile split across the two CTAs;
    covers M/N tails (a peer CTA whose 128 rows are entirely out of range) and many rounds per cluster."""
    a = _mk((K, M) if a_mn else (M, K), 11, 0.5)
    b = _mk((K, N) if b_mn else (N, K), 12, 0.5)
    bias = torch.randn(N, device="cuda")
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, act=0, impl=2)
    ref = _ref(a, b, a_mn, b_mn, bias, 0, 1.0)
    assert _rel(out, ref) < 6e-3, _rel(out, ref)
    single = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, bias=bias, act=0, impl=3)
    assert torch.equal(out, single)          # same K order, same epilogue: bit-identical to the single-CTA kernel


def test_gemm_cta_pair_gelu_and_f32_accumulate(ops):
    M, N, K = 1500, 5120, 1280
    a, b = _mk((M, K), 13, 0.3), _mk((N, K), 14, 0.3)
    bias = torch.randn(N, device="cuda") * 0.1
    out = ops.gemm(a, b, bias=bias, act=1, impl=2)
    assert _rel(out, _ref(a, b, False, False, bias, 1, 1.0)) < 8e-3
    M, N, K = 1280, 1288, 4096
    a, b = _mk((K, M), 15, 0.2), _mk((K, N), 16, 0.2)
    base = torch.randn(M, N, device="cuda")
    acc = base.clone()
    ops.gemm(a, b, a_mn=True, b_mn=True, out=acc, alpha=0.5, accumulate=True, impl=2)
    assert _rel(acc, base + _ref(a, b, True, True, None, 0, 0.5)) < 2e-5
