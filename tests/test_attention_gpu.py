"""tcgen05 flash attention (csrc/attention.cu) against torch fp32 softmax attention on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, B, T, H):
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)  # [3][B,H,T,64]
    att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1) @ v
    return att.transpose(1, 2).reshape(B * T, H * 64)


@pytest.mark.parametrize("B,T,H,scale", [(1, 128, 1, 1.0), (1, 257, 2, 1.0), (2, 1370, 12, 1.0), (1, 1370, 12, 4.0),
                                         (1, 100, 3, 2.0)])
def test_attention_matches_fp32(cuda_device, B, T, H, scale):
    from cognitive_aim_depth_estimation_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = (torch.randn(B * T, 3 * H * 64, generator=g) * scale).to(cuda_device).bfloat16()
    out = torch.full((B * T, H * 64), float("nan"), device=cuda_device, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, T, H)
    torch.cuda.synchronize()
    ref = _ref(qkv, B, T, H)
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 8e-3, (rel, err)


def test_attention_reference_jumps_mid_sequence(cuda_device):
    """Keys whose scores dwarf everything seen so far arrive late in the sequence (x 40 from token 700 on, x 0.02 again
    from 1000 on): the running reference of the softmax must be raised — lazily, at the next step, or at once with the
    step redone when it is off by more than 2^24 — without losing the earlier contributions."""
    from cognitive_aim_depth_estimation_b200 import ops
    B, T, H = 1, 1370, 4
    g = torch.Generator().manual_seed(99)
    qkv = torch.randn(B * T, 3 * H * 64, generator=g)
    k = qkv[:, H * 64: 2 * H * 64]
    k[700:1000] *= 40.0
    k[1000:] *= 0.02
    qkv = qkv.to(cuda_device).bfloat16()
    out = torch.full((B * T, H * 64), float("nan"), device=cuda_device, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, T, H)
    torch.cuda.synchronize()
    ref = _ref(qkv, B, T, H)
    assert torch.isfinite(out.float()).all()
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    assert rel < 8e-3, rel


def test_captured_and_eager_launches_on_two_streams_do_not_share_a_work_counter(cuda_device):
    """A launch recorded into a CUDA graph keeps its work-counter slot for the life of the graph; eager launches recycle a
    separate ring.  Replay the graph on one stream while > 1024 eager launches (a full turn of that ring) run on
    another: both results must stay exact."""
    from cognitive_aim_depth_estimation_b200 import ops
    B, T, H = 2, 333, 12
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv1 = (torch.randn(B * T, 3 * H * 64, generator=g) * 0.5).to(cuda_device).bfloat16()
    qkv2 = (torch.randn(B * T, 3 * H * 64, generator=g) * 0.5).to(cuda_device).bfloat16()
    ref1, ref2 = torch.empty(B * T, H * 64, device=cuda_device, dtype=torch.bfloat16), torch.empty(B * T, H * 64, device=cuda_device, dtype=torch.bfloat16)
    ops.attention(qkv1, ref1, B, T, H)
    ops.attention(qkv2, ref2, B, T, H)
    torch.cuda.synchronize()
    out1, out2 = torch.zeros_like(ref1), torch.zeros_like(ref2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s1):
        with torch.cuda.graph(graph, stream=s1):
            for _ in range(4):
                ops.attention(qkv1, out1, B, T, H)
    torch.cuda.synchronize()
    for rep in range(6):
        out1.zero_()
        out2.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            for _ in range(20):
                graph.replay()
        with torch.cuda.stream(s2):
            for _ in range(200):
                ops.attention(qkv2, out2, B, T, H)
        torch.cuda.synchronize()
        assert torch.equal(out1, ref1), rep
        assert torch.equal(out2, ref2), rep
