"""tcgen05 GEMM (csrc/gemm.cu) against torch fp32 matmul of the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def _relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 768), (300, 768, 768), (1370 * 2, 2304, 768),
                                   (1000, 768, 3072), (2738, 768, 592)])
def test_gemm_f32(cuda_device, M, N, K):
    from cognitive_aim_depth_estimation_b200 import ops
    A = _rand((M, K), cuda_device, 1).bfloat16()
    W = _rand((N, K), cuda_device, 2, 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), device=cuda_device)
    ops.gemm(A, W, ops.EPI_F32, out)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    assert torch.isfinite(out).all()
    assert _relerr(out, ref) < 2e-5, _relerr(out, ref)


def test_gemm_bias_bf16_and_gelu(cuda_device):
    from cognitive_aim_depth_estimation_b200 import ops
    M, N, K = 1370, 3072, 768
    A = _rand((M, K), cuda_device, 3).bfloat16()
    W = _rand((N, K), cuda_device, 4, 0.04).bfloat16()
    bias = _rand((N,), cuda_device, 5, 0.1)
    ref = A.float() @ W.float().t() + bias
    out = torch.zeros((M, N), device=cuda_device, dtype=torch.bfloat16)
    ops.gemm(A, W, ops.EPI_BIAS_BF16, out, bias=bias)
    assert _relerr(out, ref) < 4e-3
    ops.gemm(A, W, ops.EPI_GELU_BF16, out, bias=bias)
    assert _relerr(out, torch.nn.functional.gelu(ref)) < 4e-3


def test_gelu_epilogue_pointwise_accuracy(cuda_device):
    """The MUFU-free erf-GELU of the fc1 epilogue (odd minimax polynomial + saturating FMA) against torch's erf GELU on
    a dense sweep of pre-activations in [-12, 12]: the pre-activation is delivered exactly through the bias of a
    zero-weight GEMM, so the only error measured is the GELU approximation plus the bf16 rounding of the output."""
    from cognitive_aim_depth_estimation_b200 import ops
    M, N, K = 128, 3072, 64
    A = torch.zeros((M, K), device=cuda_device, dtype=torch.bfloat16)
    W = torch.zeros((N, K), device=cuda_device, dtype=torch.bfloat16)
    for lo, hi in ((-12.0, 12.0), (-5.0, 5.0), (-1.0, 1.0), (-0.01, 0.01)):
        bias = torch.linspace(lo, hi, N, device=cuda_device)
        out = torch.zeros((M, N), device=cuda_device, dtype=torch.bfloat16)
        ops.gemm(A, W, ops.EPI_GELU_BF16, out, bias=bias)
        torch.cuda.synchronize()
        ref = torch.nn.functional.gelu(bias.double()).float()
        got = out[0].float()
        assert torch.equal(out[0], out[-1])
        # bf16 rounding of the exact value is the floor: half an ulp = 2^-9 relative
        err = (got - ref).abs()
        assert (err <= ref.abs() * 2.0 ** -8 + 6e-5).all(), (err - ref.abs() * 2.0 ** -8).max()


def test_gemm_resid_inplace(cuda_device):
    from cognitive_aim_depth_estimation_b200 import ops
    M, N, K = 1370 * 3, 768, 3072
    A = _rand((M, K), cuda_device, 6).bfloat16()
    W = _rand((N, K), cuda_device, 7, 0.02).bfloat16()
    bias = _rand((N,), cuda_device, 8, 0.1)
    ls = _rand((N,), cuda_device, 9, 1.0)
    x = _rand((M, N), cuda_device, 10)
    ref = x + ls * (A.float() @ W.float().t() + bias)
    ops.gemm(A, W, ops.EPI_RESID_F32, x, bias=bias, ls=ls)
    assert _relerr(x, ref) < 1e-5


def test_gemm_patch_embed(cuda_device):
    from cognitive_aim_depth_estimation_b200 import ops
    B, Np, D, K = 3, 256, 768, 592
    A = _rand((B * Np, K), cuda_device, 11).bfloat16()
    A[:, 588:] = 0
    W = _rand((D, K), cuda_device, 12, 0.05).bfloat16()
    bias = _rand((D,), cuda_device, 13, 0.1)
    pos = _rand((Np + 1, D), cuda_device, 14, 0.02)
    x = torch.zeros((B, Np + 1, D), device=cuda_device)
    ops.gemm(A, W, ops.EPI_PATCH_F32, x.view(-1, D), bias=bias, pos=pos, patches_per_img=Np)
    ref = (A.float() @ W.float().t() + bias).view(B, Np, D) + pos[1:]
    assert _relerr(x[:, 1:], ref) < 1e-5
    assert (x[:, 0] == 0).all()


@pytest.mark.parametrize("N,D", [(256, 768), (1369, 768)])
def test_gemm_focal_stats(cuda_device, N, D):
    """Row statistics (pass A) and softmax column sums (pass B) == column-mean of a row softmax."""
    from cognitive_aim_depth_estimation_b200 import ops
    B = 2
    q = _rand((B, N, D), cuda_device, 15).bfloat16()
    k = _rand((B, N, D), cuda_device, 16, 0.5).bfloat16()
    scale = 1.0 / math.sqrt(96.0)
    P = ops.stats_partials(N)
    pm = torch.zeros((B, P, N), device=cuda_device)   # span-major partials
    ps = torch.zeros((B, P, N), device=cuda_device)
    ops.gemm(q, k, ops.EPI_ROWSTATS, None, M=N, N=N, K=D, lda=D, ldw=D, batch=B, a_batch_stride=N * D,
             w_batch_stride=N * D, scale_log2=scale * ops.LOG2E, part_a=pm, part_b=ps)
    s = torch.einsum("bid,bjd->bij", q.float(), k.float()) * scale
    s2 = s * ops.LOG2E
    rmax = pm.max(dim=1).values
    assert torch.allclose(rmax, s2.max(dim=-1).values, rtol=1e-5, atol=1e-4)
    rsum = (ps * torch.exp2(pm - rmax[:, None, :])).sum(1)
    ref_sum = torch.exp2(s2 - s2.max(dim=-1, keepdim=True).values).sum(-1)
    assert torch.allclose(rsum, ref_sum, rtol=2e-3)
    # pass B: A = keys, W = queries  ->  acc[j, i] = s[i, j]
    pc = torch.zeros((B, P, N), device=cuda_device)
    rinv = (1.0 / rsum).contiguous()
    ops.gemm(k, q, ops.EPI_COLSUM, None, M=N, N=N, K=D, lda=D, ldw=D, batch=B, a_batch_stride=N * D,
             w_batch_stride=N * D, scale_log2=scale * ops.LOG2E, part_a=pc, col_max=rmax.contiguous(), col_rinv=rinv)
    colmean = pc.sum(1) / N
    ref = torch.softmax(s, dim=-1).mean(dim=1)
    assert torch.allclose(colmean, ref, rtol=5e-3, atol=1e-7), (colmean - ref).abs().max()


@pytest.mark.parametrize("N,D", [(256, 768), (1369, 768), (100, 768)])
def test_focal_colsum_from_stored_exponentials(cuda_device, N, D):
    """Pass A keeps exp2(s - span max) as fp16; ca_colsum_e turns it into the column sums of the row softmax
    (== column mean of src/model.py:234 up to the 1/N), un-weighted and weighted (un-guided value path, :308)."""
    from cognitive_aim_depth_estimation_b200 import ops
    B = 2
    q = _rand((B, N, D), cuda_device, 21).bfloat16()
    k = _rand((B, N, D), cuda_device, 22, 0.5).bfloat16()
    scale = 1.0 / math.sqrt(96.0)
    P = ops.stats_partials(N)
    pm = torch.zeros((B, P, N), device=cuda_device)   # span-major partials
    ps = torch.zeros((B, P, N), device=cuda_device)
    lde = 64 * ((N + 63) // 64)
    E = torch.full((B, N, lde), float("nan"), device=cuda_device, dtype=torch.float16)
    ops.gemm(q, k, ops.EPI_ROWSTATS, E, M=N, N=N, K=D, lda=D, ldw=D, batch=B, a_batch_stride=N * D,
             w_batch_stride=N * D, scale_log2=scale * ops.LOG2E, part_a=pm, part_b=ps, ldo=lde, out_batch_stride=N * lde)
    assert torch.isfinite(E.float()).all() and float(E.max()) <= 1.0 and (E[:, :, N:] == 0).all()
    s = torch.einsum("bid,bjd->bij", q.float(), k.float()) * scale
    ref = torch.softmax(s, dim=-1)
    for weight in (None, torch.rand(B, N, device=cuda_device)):
        wtab = torch.empty((B, P, N), device=cuda_device)
        pc = torch.zeros((B, P, N), device=cuda_device)   # span-major partials
        ops.rowstats_merge(pm, ps, weight, None, None, wtab)
        ops.colsum_e(E, wtab, pc, B, N)
        got = pc.sum(1)
        want = ref.sum(dim=1) if weight is None else (ref * weight[:, :, None]).sum(dim=1)
        assert torch.allclose(got, want, rtol=3e-3, atol=1e-6), ((got - want).abs() / want).max()


def _stats_to_mean_var(stats):
    """[M, slots, 2] (sum, M2) per 128-column span -> (mean, biased variance) per row (Chan's combination)."""
    s, m2 = stats[..., 0].double(), stats[..., 1].double()
    slots = stats.shape[1]
    mean = s.sum(1) / (128 * slots)
    var = (m2.sum(1) + (128 * (s / 128 - mean[:, None]) ** 2).sum(1)) / (128 * slots)
    return mean, var


@pytest.mark.parametrize("M", [8, 1370, 2741])
def test_ln_shadow_rows_and_statistics(cuda_device, M):
    from cognitive_aim_depth_estimation_b200 import ops
    x = _rand((M, 768), cuda_device, 30, 3.0) + _rand((M, 1), cuda_device, 31, 2.0)
    shadow = torch.zeros(M, 768, device=cuda_device, dtype=torch.bfloat16)
    stats = torch.full((M, 6, 2), float("nan"), device=cuda_device)
    ops.ln_shadow(x, shadow, stats)
    torch.cuda.synchronize()
    assert torch.equal(shadow, x.bfloat16())
    mean, var = _stats_to_mean_var(stats)
    assert torch.allclose(mean, x.double().mean(1), atol=1e-5)
    assert torch.allclose(var, x.double().var(1, unbiased=False), rtol=1e-5)


@pytest.mark.parametrize("M,K", [(40, 768), (1370, 768), (2741, 768), (1370 * 3, 3072), (300, 832)])
def test_gemm_resid_ln_updates_rows_and_leaves_shadow_and_statistics(cuda_device, M, K):
    """EPI_RESID_LN_F32: the residual update of EPI_RESID_F32 with the old rows read into the SM (TMA load, update in
    shared memory, TMA store), plus what the LayerNorm after it needs: the new rows as bf16 and per-span (sum, M2)."""
    from cognitive_aim_depth_estimation_b200 import ops
    N = 768
    A = _rand((M, K), cuda_device, 32).bfloat16()
    W = _rand((N, K), cuda_device, 33, 0.03).bfloat16()
    bias = _rand((N,), cuda_device, 34, 0.1)
    ls = _rand((N,), cuda_device, 35, 1.0)
    x0 = _rand((M, N), cuda_device, 36, 2.0) + 1.5
    ref = x0 + ls * (A.float() @ W.float().t() + bias)
    x = x0.clone()
    shadow = torch.zeros(M, N, device=cuda_device, dtype=torch.bfloat16)
    stats = torch.full((M, 6, 2), float("nan"), device=cuda_device)
    ops.gemm_ln(A, W, ops.EPI_RESID_LN_F32, x, bias=bias, ls=ls, stats=stats, shadow=shadow)
    torch.cuda.synchronize()
    assert torch.isfinite(x).all() and torch.isfinite(stats).all()
    assert _relerr(x, ref) < 2e-5, _relerr(x, ref)
    assert torch.equal(shadow, x.bfloat16())
    mean, var = _stats_to_mean_var(stats)
    assert torch.allclose(mean, x.double().mean(1), atol=2e-5)
    assert torch.allclose(var, x.double().var(1, unbiased=False), rtol=2e-5)
    # the plain reduce-add epilogue computes the same update
    x2 = x0.clone()
    ops.gemm(A, W, ops.EPI_RESID_F32, x2, bias=bias, ls=ls)
    assert _relerr(x2, x) < 1e-6


@pytest.mark.parametrize("M,N,gelu", [(1370, 2304, False), (2741, 3072, True), (40, 2304, False)])
def test_gemm_with_folded_layernorm_matches_layernorm_then_gemm(cuda_device, M, N, gelu):
    """EPI_LN_*: LayerNorm(x) W^T + b computed as rstd * (bf16(x) W'^T) + b' (row-centred W') from the raw rows and the
    statistics buffer, against torch's LayerNorm + matmul in fp32, with a row mean comparable to the row deviation."""
    from cognitive_aim_depth_estimation_b200 import ops
    K = 768
    x = _rand((M, K), cuda_device, 40, 2.0) + _rand((M, 1), cuda_device, 41, 1.0)
    gamma = 1.0 + _rand((K,), cuda_device, 42, 0.2)
    beta = _rand((K,), cuda_device, 43, 0.1)
    W = _rand((N, K), cuda_device, 44, 0.03)
    bias = _rand((N,), cuda_device, 45, 0.1)
    ref = torch.nn.functional.layer_norm(x, (K,), gamma, beta, eps=1e-6) @ W.t() + bias
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    shadow = torch.zeros(M, K, device=cuda_device, dtype=torch.bfloat16)
    stats = torch.zeros(M, 6, 2, device=cuda_device)
    ops.ln_shadow(x, shadow, stats)
    wp, bp = ops.fold_layernorm(W, bias, gamma, beta)
    out = torch.zeros(M, N, device=cuda_device, dtype=torch.bfloat16)
    ops.gemm_ln(shadow, wp, ops.EPI_LN_GELU_BF16 if gelu else ops.EPI_LN_BIAS_BF16, out, bias=bp, stats=stats)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    # the unfused path (LayerNorm kernel -> bf16 -> GEMM) on the same inputs, for scale
    h = torch.zeros(M, K, device=cuda_device, dtype=torch.bfloat16)
    ops.layernorm(x, gamma, beta, h)
    out2 = torch.zeros(M, N, device=cuda_device, dtype=torch.bfloat16)
    ops.gemm(h, W.bfloat16(), ops.EPI_GELU_BF16 if gelu else ops.EPI_BIAS_BF16, out2, bias=bias)
    e_fold, e_plain = _relerr(out, ref), _relerr(out2, ref)
    assert e_fold < 6e-3, (e_fold, e_plain)
    assert e_fold < 2.0 * e_plain + 1e-3, (e_fold, e_plain)
