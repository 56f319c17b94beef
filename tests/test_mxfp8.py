"""MXFP8 block-scaled GEMM: format oracle on the CPU, tcgen05 kernels against it on the GPU."""
import pytest
import torch

from hfl_b200.ops import mxfp8 as mx


def test_mx_format_oracle_roundtrip():
    torch.manual_seed(0)
    x = (torch.randn(200, 256) * torch.logspace(-3, 3, 200).view(-1, 1)).bfloat16()
    x[5, 32:64] = 0.0                                     # an all-zero block
    q, sf = mx.quantize(x)
    assert q.dtype == torch.uint8 and sf.numel() == 2 * 2 * 512
    d = mx.dequantize(q, sf)
    err = (d - x.float()).abs()
    blk_amax = x.float().view(200, 8, 32).abs().amax(2, keepdim=True).expand(200, 8, 32).reshape(200, 256)
    assert float((err / blk_amax.clamp_min(1e-30)).max()) <= 2 ** -3      # e4m3: 3 mantissa bits, block-relative
    assert float(d[5, 32:64].abs().sum()) == 0.0
    a, b = torch.randn(64, 256).bfloat16(), torch.randn(96, 256).bfloat16()
    c = mx.mx_linear(a, b)
    ref = a.float() @ b.float().t()
    assert float((c.float() - ref).norm() / ref.norm()) < 0.06


@pytest.mark.gpu
@pytest.mark.parametrize("R,K", [(128, 128), (300, 256), (77, 1024)])
def test_mx_quantize_kernel_matches_oracle(R, K):
    from hfl_b200.ops import cuda as cu
    assert cu.available()
    dev = torch.device("cuda")
    torch.manual_seed(1)
    x = (torch.randn(R, K, device=dev) * torch.logspace(-4, 4, R, device=dev).view(-1, 1)).bfloat16()
    x[0, :32] = 0.0
    q, sf = mx.quantize(x)
    qr, er = mx.quantize_ref(x)
    assert torch.equal(q.view(torch.float8_e4m3fn).float(), qr.float())
    got_e = sf[mx.sf_offsets(R, K, dev)].to(torch.int32) - 127
    assert torch.equal(got_e, er)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (128, 128, 512), (256, 384, 256), (200, 130, 1024), (6400, 1024, 256)])
def test_mxfp8_gemm_matches_dequantised_reference(M, N, K):
    """The tensor-core result must equal the fp32 product of the DEQUANTISED operands (exact products, fp32 sums)."""
    from hfl_b200.ops import cuda as cu
    assert cu.available()
    dev = torch.device("cuda")
    torch.manual_seed(2)
    a = (torch.randn(M, K, device=dev) * (1 + 3 * torch.rand(M, 1, device=dev))).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    aq, sfa = mx.quantize(a)
    bq, sfb = mx.quantize(b)
    ref = mx.dequantize(aq, sfa) @ mx.dequantize(bq, sfb).t()
    c = mx.mx_gemm(aq, sfa, bq, sfb, None, 1.0, out_fp32=True)
    torch.cuda.synchronize()
    assert torch.allclose(c, ref, rtol=2e-3, atol=2e-3 * float(ref.abs().max())), float((c - ref).abs().max())
    c2 = mx.mx_gemm(aq, sfa, bq, sfb, bias, 0.5, out_fp32=False)
    assert torch.allclose(c2.float(), 0.5 * ref + bias, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
    # end to end against the unquantised product: the format's error budget
    full = a.float() @ b.float().t()
    rel = float((c - full).norm() / full.norm())
    print("MXFP8 %dx%dx%d: relative error vs bf16 operands %.4f" % (M, N, K, rel))
    assert rel < 0.06
