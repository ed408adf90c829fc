"""The handle-level C-ABI (ca_create / ca_forward_guided / ca_forward / ca_backbone / ca_destroy, include/cogaim_b200.h;
SURVEY.md §8b, VERDICT r1 missing #3): the whole forward behind one native call, driven through
cognitive_aim_depth_estimation_b200/native.py — which does NOT use model.py — and held to the same parity bar against the
CPU oracle (depth abs-rel 1e-2, confidence / heat-map 1e-2, arg-max cell exact)."""
import sys

import pytest
import torch

from oracle import cogaim_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    return orc.build_state_dict(0)


@pytest.fixture(scope="module")
def native(cuda_device, sd):
    from cognitive_aim_depth_estimation_b200.native import NativeModel
    m = NativeModel(sd, device=cuda_device)
    yield m
    m.close()


def _exif(ex):
    return {k: v.cuda() for k, v in ex.items()}


def _check(depth, conf, heat, ref):
    abs_rel = ((depth.cpu() - ref["depth"]).abs() / ref["depth"].abs()).max().item()
    assert abs_rel <= 1e-2, abs_rel
    assert (conf.cpu() - ref["confidence"]).abs().max().item() <= 1e-2
    assert (heat.cpu() - ref["heatmap"]).abs().max().item() <= 1e-2
    assert torch.equal(heat.cpu().argmax(-1), ref["heatmap"].argmax(-1))


def test_native_module_does_not_use_model_py():
    import cognitive_aim_depth_estimation_b200.native as nat
    src = open(nat.__file__).read()
    assert "from .model" not in src and "import model" not in src and "oracle" not in src.replace("oracle)", "")


@pytest.mark.parametrize("S", [224, 518])
def test_guided_forward_through_the_handle(native, sd, S):
    x, ex = orc.synthetic_images(2, S), orc.synthetic_exif(2)
    tokens = orc.dinov2_tokens(sd, x)
    for instruction in ("center", "top-left", "bottomright", "somewhere else"):
        torch.manual_seed(11)
        ref = orc.forward_with_guidance(sd, None, ex, instruction, tokens=tokens, update_history=False)
        # eager call, graph capture, graph replay (a non-default stream: the legacy stream cannot be captured, callers on
        # it get eager launches): all must agree with the oracle
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        for i in range(4):
            with torch.cuda.stream(side if i else torch.cuda.current_stream()):
                torch.manual_seed(11)
                depth, conf, heat, arg = native.forward_with_guidance(x.cuda(), _exif(ex), instruction)
                torch.cuda.synchronize()
            _check(depth, conf, heat, ref)
            assert torch.equal(arg.cpu().long(), ref["heatmap"].argmax(-1))
    assert native.launch_count() > 80
    # an explicit guidance tensor of another grid size is the caller's to resize; one of the right size is taken as is
    g = S // 14
    guide = torch.rand(g * g, generator=torch.Generator().manual_seed(3)) * 3
    torch.manual_seed(11)
    ref = orc.forward_with_guidance(sd, None, ex, guide, tokens=tokens, update_history=False)
    torch.manual_seed(11)
    depth, conf, heat, _ = native.forward_with_guidance(x.cuda(), _exif(ex), guide)
    _check(depth, conf, heat, ref)


def test_unguided_forward_and_backbone_through_the_handle(native, sd):
    x, ex = orc.synthetic_images(2, 224), orc.synthetic_exif(2)
    tokens = orc.dinov2_tokens(sd, x)
    tok = native.backbone_tokens(x.cuda()).cpu()
    assert ((tok - tokens).norm() / tokens.norm()).item() < 1.5e-2
    for with_exif in (True, False):
        ref = orc.forward_unguided(sd, None, ex if with_exif else None, tokens=tokens, update_history=False)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        for _ in range(3):
            with torch.cuda.stream(side):
                depth, conf, att, fused = native.forward(x.cuda(), _exif(ex) if with_exif else None, runs=2)
                torch.cuda.synchronize()
            _check(depth, conf, att, ref)
            f = ref["fusion_features"]
            assert ((fused.cpu() - f).norm() / f.norm()).item() < 1e-2


def test_handle_matches_the_python_orchestration(native, sd, cuda_device):
    """Same kernels, same order: the native launcher and model.py agree to rounding of the host-built tables (C++ sinf /
    expf vs torch's), and both keep the CuriosityModule's ring buffer like the reference (src/model.py:760-773)."""
    from cognitive_aim_depth_estimation_b200.model import create_model
    from cognitive_aim_depth_estimation_b200.native import NativeModel
    m = create_model({"model": {}}, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    fresh = NativeModel(sd, device=cuda_device, fold_layernorm=m._ln_folded())  # same kernels, same operands
    x, ex = orc.synthetic_images(3, 518, seed=5), orc.synthetic_exif(3, seed=6)
    for instruction in ("left", "bottom"):
        torch.manual_seed(21)
        a = m.forward_with_guidance(x.cuda(), _exif(ex), instruction, return_attention=True)
        torch.manual_seed(21)
        b = fresh.forward_with_guidance(x.cuda(), _exif(ex), instruction)
        for p, q in zip(a, b[:3]):
            assert torch.allclose(p, q, rtol=1e-4, atol=1e-6), (p - q).abs().max()
        assert torch.equal(m._last_argmax, b[3])
    torch.cuda.synchronize()
    assert int(fresh.history_pointer) == int(m.curiosity_module.history_pointer) == 6
    assert torch.allclose(fresh.exploration_history[:6], m.curiosity_module.exploration_history[:6], rtol=1e-4, atol=1e-7)
    fresh.close()


def test_handle_with_caller_folded_layernorm(sd, cuda_device):
    """ca_model_weights.layernorm_folded = 1: the caller folds norm1 / norm2 into the q/k/v and fc1 operands and the handle
    runs the encoder without LayerNorm passes (24 launches fewer); same tolerances against the oracle."""
    from cognitive_aim_depth_estimation_b200.native import NativeModel
    plain, folded = NativeModel(sd, device=cuda_device), NativeModel(sd, device=cuda_device, fold_layernorm=True)
    x = orc.synthetic_images(2, 224)
    ref = orc.dinov2_tokens(sd, x)
    a, b = plain.backbone_tokens(x.cuda()).cpu(), folded.backbone_tokens(x.cuda()).cpu()
    n_plain, n_folded = plain.launch_count(), folded.launch_count()
    assert ((a - ref).norm() / ref.norm()).item() < 1.5e-2
    assert ((b - ref).norm() / ref.norm()).item() < 1.5e-2
    assert n_folded == n_plain - 24 + 1, (n_plain, n_folded)   # 24 LayerNorm launches gone, one ca_ln_shadow added
    plain.close()
    folded.close()


def test_handle_errors_are_reported(native):
    from cognitive_aim_depth_estimation_b200._lib import CogAimError
    x, ex = orc.synthetic_images(1, 224), orc.synthetic_exif(1)
    with pytest.raises(CogAimError):
        native.forward_with_guidance(x[:, :, :100, :100].cuda(), _exif(ex), "center")     # side not a multiple of 14
    with pytest.raises(CogAimError):
        native.forward_with_guidance(x.cuda(), None, "center")                             # guided needs EXIF
