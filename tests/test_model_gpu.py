"""End-to-end parity of the B200 path against the CPU oracle on the same random-init weights (seed protocol of
SURVEY.md §8c): depth abs-rel <= 1e-2, confidence / heatmap max-abs <= 1e-2, heatmap argmax cell bit-exact
(BASELINE.json north_star tolerances, bf16 tensor-core operands with fp32 accumulation)."""
import numpy as np
import pytest
import torch

from oracle import cogaim_oracle as orc

pytestmark = pytest.mark.gpu

CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
DEPTH_ABS_REL, CONF_MAX_ABS, HEAT_MAX_ABS, TOKEN_REL_FRO = 1e-2, 1e-2, 1e-2, 1.5e-2


@pytest.fixture(scope="module")
def sd():
    return orc.build_state_dict(0)


@pytest.fixture(scope="module")
def model(cuda_device, sd):
    from cognitive_aim_depth_estimation_b200.model import create_model
    m = create_model(CFG, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    return m


@pytest.fixture(scope="module")
def cases(sd):
    """(S, B) -> images, exif, oracle tokens."""
    out = {}
    for S, B in ((224, 2), (518, 2)):
        x = orc.synthetic_images(B, S)
        out[(S, B)] = (x, orc.synthetic_exif(B), orc.dinov2_tokens(sd, x))
    return out


def _cuda_exif(ex, dev):
    return {k: v.to(dev) for k, v in ex.items()}


@pytest.mark.parametrize("S,B", [(224, 2), (518, 2)])
def test_backbone_tokens(model, cases, S, B):
    x, _, ref = cases[(S, B)]
    tok = model.backbone_tokens(x.cuda()).cpu()
    rel = ((tok - ref).norm() / ref.norm()).item()
    assert torch.isfinite(tok).all()
    assert rel < TOKEN_REL_FRO, rel


@pytest.mark.parametrize("S,B", [(224, 2), (518, 2)])
@pytest.mark.parametrize("instruction", orc.INSTRUCTIONS)
def test_guided_parity(model, sd, cases, S, B, instruction):
    x, ex, tokens = cases[(S, B)]
    torch.manual_seed(11)
    ref = orc.forward_with_guidance(sd, None, ex, instruction, tokens=tokens, update_history=False)
    torch.manual_seed(11)
    depth, conf, heat = model.forward_with_guidance(x.cuda(), _cuda_exif(ex, "cuda"), instruction, return_attention=True)
    depth, conf, heat = depth.cpu(), conf.cpu(), heat.cpu()
    assert depth.shape == (B, 1) and conf.shape == (B, 1) and heat.shape == ref["heatmap"].shape
    abs_rel = ((depth - ref["depth"]).abs() / ref["depth"].abs()).max().item()
    assert abs_rel <= DEPTH_ABS_REL, abs_rel
    assert (conf - ref["confidence"]).abs().max().item() <= CONF_MAX_ABS
    assert (heat - ref["heatmap"]).abs().max().item() <= HEAT_MAX_ABS
    assert torch.allclose(heat.sum(-1), torch.ones(B), atol=1e-5)
    ref_arg = ref["heatmap"].argmax(-1)
    top2 = ref["heatmap"].topk(2, dim=-1).values
    margin = ((top2[:, 0] - top2[:, 1]) / top2[:, 0]).min().item()
    assert torch.equal(heat.argmax(-1), ref_arg), f"argmax differs (oracle top-1/top-2 relative margin {margin:.2e})"
    assert torch.equal(model._last_argmax.cpu().long(), ref_arg)
    assert model.get_attention_weights() is not None


def test_focal_attention_given_oracle_tokens(model, sd, cases):
    """The last-iteration focal attention (what decides the argmax inside the top mask ring), isolated from backbone
    error by feeding the oracle's fp32 tokens: bf16 Q/K operands, fp32 statistics."""
    for key in ((224, 2), (518, 2)):
        _, _, tokens = cases[key]
        fused_ref, ref = orc.iterative_focal_stream(sd, tokens[:, 1:], need_features=True)
        att, feat = model.focal_attention(tokens.cuda(), want_features=True)
        rel = ((att.cpu() - ref).abs() / ref)
        assert rel.max().item() < 5e-2 and rel.mean().item() < 2e-3, (rel.max().item(), rel.mean().item())
        assert torch.allclose(att.sum(-1).cpu(), torch.ones(att.shape[0]), atol=1e-5)
        assert ((feat.cpu() - fused_ref).norm() / fused_ref.norm()).item() < 1e-2


@pytest.mark.parametrize("with_exif", [True, False])
def test_unguided_parity(model, sd, cases, with_exif):
    x, ex, tokens = cases[(224, 2)]
    ex = ex if with_exif else None
    ref = orc.forward_unguided(sd, None, ex, tokens=tokens, update_history=False)
    if hasattr(model, "_last_attention_weights"):
        delattr(model, "_last_attention_weights")
    depth, conf, att = model(x.cuda(), _cuda_exif(ex, "cuda") if ex else None, return_attention=True)
    abs_rel = ((depth.cpu() - ref["depth"]).abs() / ref["depth"].abs()).max().item()
    assert abs_rel <= DEPTH_ABS_REL, abs_rel
    assert (conf.cpu() - ref["confidence"]).abs().max().item() <= CONF_MAX_ABS
    assert (att.cpu() - ref["heatmap"]).abs().max().item() <= HEAT_MAX_ABS
    assert torch.equal(att.cpu().argmax(-1), ref["heatmap"].argmax(-1))
    f = model.fusion_features.cpu()
    assert ((f - ref["fusion_features"]).norm() / ref["fusion_features"].norm()).item() < 1e-2
    assert torch.equal(model.get_attention_weights().cpu(), att.cpu())


def test_confidence_path_exercised(cuda_device, sd, cases):
    """SURVEY.md §4 degeneracy guard: at seed-0 init the confidence is the constant sigmoid(2.0); flip the sign of
    confidence_head.0.weight so that ReLU passes and the second Linear + Sigmoid actually matter."""
    from cognitive_aim_depth_estimation_b200.model import create_model
    sd2 = dict(sd)
    sd2["confidence_head.0.weight"] = -sd["confidence_head.0.weight"]
    sd2["confidence_head.0.bias"] = sd["confidence_head.0.bias"] + 0.5
    m = create_model(CFG, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd2)
    x, ex, tokens = cases[(224, 2)]
    torch.manual_seed(11)
    ref = orc.forward_with_guidance(sd2, None, ex, "left", tokens=tokens, update_history=False)
    torch.manual_seed(11)
    _, conf = m.forward_with_guidance(x.cuda(), _cuda_exif(ex, "cuda"), "left")
    assert not np.allclose(ref["confidence"].numpy(), 0.8807970285, atol=1e-4)
    assert (conf.cpu() - ref["confidence"]).abs().max().item() <= CONF_MAX_ABS


def test_guided_without_exif_falls_back_like_reference(model, sd, cases):
    """exif_data=None in guided mode: the reference stores the guided heatmap, fails in `fusion`, and returns the
    un-guided forward (src/model.py:1212,1237-1240)."""
    x, ex, tokens = cases[(224, 2)]
    ref_u = orc.forward_unguided(sd, None, None, tokens=tokens, update_history=False)
    torch.manual_seed(11)
    ref_g = orc.forward_with_guidance(sd, None, ex, "top", tokens=tokens, update_history=False)
    depth, conf, att = model.forward_with_guidance(x.cuda(), None, "top", return_attention=True)
    assert ((depth.cpu() - ref_u["depth"]).abs() / ref_u["depth"]).max().item() <= DEPTH_ABS_REL
    assert (att.cpu() - ref_u["heatmap"]).abs().max().item() <= HEAT_MAX_ABS
    assert torch.equal(model.get_attention_weights().cpu().argmax(-1), ref_g["heatmap"].argmax(-1))


def test_uint8_preprocess_path(model, cases):
    """uint8 HWC -> fused normalise+patchify kernel == float path on the ToTensor/Normalize result (demo.py:162-166)."""
    from cognitive_aim_depth_estimation_b200 import ops
    u8 = torch.randint(0, 256, (2, 224, 224, 3), generator=torch.Generator().manual_seed(1235), dtype=torch.uint8)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    xf = ((u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std).cuda()
    t_float = model.backbone_tokens(xf).clone()
    t_u8 = model.tokens_from_uint8(u8.cuda())
    assert ((t_u8 - t_float).norm() / t_float.norm()).item() < 2e-3


def test_errors_are_raised_not_swallowed(model):
    with pytest.raises(ValueError):
        model.forward_with_guidance(torch.zeros(1, 3, 224, 224).cuda(), {"focal_length": torch.ones(1)}, "center")
    with pytest.raises(ValueError):
        model(torch.zeros(1, 1, 224, 224).cuda())
    bad = {"focal_length": torch.ones(1), "aperture": torch.ones(1), "iso": torch.ones(1),
           "camera_idx": torch.tensor([99])}
    with pytest.raises(ValueError):
        model.forward_with_guidance(torch.zeros(1, 3, 224, 224).cuda(), bad, "center")


def _lora_sd(scale=0.05):
    sd = orc.build_state_dict(0, use_lora=True)
    g = torch.Generator().manual_seed(77)
    for i in range(12):
        sd[f"lora_layers.{i}.lora_B"] = torch.randn(768, 16, generator=g) * scale
    return sd


def test_lora_adapters_are_ignored_like_the_reference(cuda_device):
    """`use_lora: true` with NON-zero lora_B: the reference never applies its adapters (src/model.py:30, 824-831), so the
    outputs must equal the fixture the unmodified reference produced from these very weights (tests/golden/lora.npz)."""
    import os
    from cognitive_aim_depth_estimation_b200.model import create_model
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "lora.npz"))
    m = create_model(dict(CFG, use_lora=True), {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(_lora_sd())
    torch.manual_seed(11)
    d, c, h = m.forward_with_guidance(orc.synthetic_images(2, 224).cuda(), _cuda_exif(orc.synthetic_exif(2), "cuda"),
                                      "center", return_attention=True)
    assert (np.abs(d.cpu().numpy() - gold["depth"]) / np.abs(gold["depth"])).max() <= DEPTH_ABS_REL
    assert np.abs(h.cpu().numpy() - gold["heat"]).max() <= HEAT_MAX_ABS
    assert (h.cpu().numpy().argmax(-1) == gold["heat"].argmax(-1)).all()


@pytest.mark.parametrize("mode", ["merge", "fused"])
@pytest.mark.parametrize("target", ["query", "value", "attention_output"])
def test_lora_merge_target_applies_the_adapters(cuda_device, target, mode):
    """The opt-in (non-reference) `lora_merge_target`: y = W x + (alpha / rank) B A x on the chosen 768 -> 768 projection
    of every encoder layer, checked against a plain fp32 restatement (the oracle backbone with W + (16/16) B A) — with the
    adapters merged into the packed weight (`lora_mode: merge`, default) and as the fused low-rank update of the GEMM
    (`lora_mode: fused`: t = x A^T written behind each activation row, [W | s B] with one extra K block)."""
    from cognitive_aim_depth_estimation_b200.model import create_model
    sd = _lora_sd(0.5)  # B A of the same size as the weight it adapts
    name = {"query": "attention.attention.query.weight", "value": "attention.attention.value.weight",
            "attention_output": "attention.output.dense.weight"}[target]
    merged = dict(sd)
    for i in range(12):
        k = f"backbone.encoder.layer.{i}.{name}"
        merged[k] = sd[k] + sd[f"lora_layers.{i}.lora_B"] @ sd[f"lora_layers.{i}.lora_A"]
    x = orc.synthetic_images(2, 224)
    want, plain = orc.dinov2_tokens(merged, x), orc.dinov2_tokens(sd, x)
    m = create_model(dict(CFG, use_lora=True, lora_merge_target=target, lora_mode=mode), {"num_cameras": 71},
                     device=cuda_device)
    m.load_state_dict(sd)
    tok = m.backbone_tokens(x.cuda()).cpu()
    rel = ((tok - want).norm() / want.norm()).item()
    moved = ((plain - want).norm() / want.norm()).item()
    assert rel < TOKEN_REL_FRO, rel
    assert moved > 3 * rel, (moved, rel)  # the adapters really changed the tokens
    if mode == "fused":
        # per-call adapters: zero adapters give the plain backbone, the original ones give `want` again — no re-pack of W,
        # and through the already captured CUDA graph
        zero = {i: (torch.zeros(16, 768), torch.zeros(768, 16)) for i in range(12)}
        m.set_lora_adapters(zero)
        tok0 = m.backbone_tokens(x.cuda()).cpu()
        assert ((tok0 - plain).norm() / plain.norm()).item() < TOKEN_REL_FRO
        m.set_lora_adapters({i: (sd[f"lora_layers.{i}.lora_A"], sd[f"lora_layers.{i}.lora_B"]) for i in range(12)})
        tok1 = m.backbone_tokens(x.cuda()).cpu()
        assert torch.equal(tok1, tok)
    else:
        with pytest.raises(ValueError):
            m.set_lora_adapters({})


def test_camera_idx_range_is_checked_on_the_device(cuda_device, sd, cases):
    """the default API path must not synchronise.  A camera index outside the embedding table
    (nn.Embedding would raise, src/model.py:491) is clamped by the heads kernel and flagged in a pinned word: the call
    itself returns without a device-to-host sync, `check_inputs()` (or the next forward) raises ValueError."""
    from cognitive_aim_depth_estimation_b200.model import create_model
    m = create_model(CFG, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    x, ex, _ = cases[(224, 2)]
    good = _cuda_exif(ex, "cuda")
    bad = dict(good, camera_idx=torch.tensor([int(ex["camera_idx"][0]), 71], device="cuda"))
    assert m.validate_inputs
    torch.manual_seed(11)
    d0, _ = m.forward_with_guidance(x.cuda(), good, "center")
    m.check_inputs()                                   # nothing to report
    torch.manual_seed(11)
    d1, _ = m.forward_with_guidance(x.cuda(), bad, "center")   # returns: the fault is only flagged
    assert torch.isfinite(d1).all() and torch.equal(d1[0], d0[0])
    with pytest.raises(ValueError, match="camera_idx out of range"):
        m.check_inputs()
    m.check_inputs()                                   # reported once, then cleared
    m.forward_with_guidance(x.cuda(), dict(good, camera_idx=torch.tensor([-1, 0], device="cuda")), "center")
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="camera_idx out of range"):
        m.forward_with_guidance(x.cuda(), good, "center")      # ... or raised by the next call
    m.validate_inputs = False
    m.forward_with_guidance(x.cuda(), bad, "center")
    m.check_inputs()                                   # validation off: clamped silently


def test_model_on_a_non_current_device(sd, cases):
    """every launch must go to the model's own GPU, whatever the caller's current device is."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from cognitive_aim_depth_estimation_b200.model import create_model
    x, ex, _ = cases[(224, 2)]
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = create_model(CFG, {"num_cameras": 71}, device=dev)
        m.load_state_dict(sd)
        torch.cuda.set_device(0)                       # current device stays 0 for both models
        torch.manual_seed(11)
        outs.append([t.cpu() for t in m.forward_with_guidance(x.to(dev), _cuda_exif(ex, dev), "left",
                                                               return_attention=True)])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_uint8_images_through_the_forward(model, cases):
    """demo-shaped flow: uint8 [B, H, W, 3] straight into forward_with_guidance / forward (Resize when `input_size` is
    set, ToTensor, Normalize fused with the patchify) == the float path on model.preprocess() of the same images."""
    u8 = torch.randint(0, 256, (2, 300, 400, 3), generator=torch.Generator().manual_seed(5), dtype=torch.uint8).cuda()
    _, ex, _ = cases[(224, 2)]
    exg = _cuda_exif(ex, "cuda")
    model.input_size = 224
    try:
        torch.manual_seed(11)
        a = [t.clone() for t in model.forward_with_guidance(u8, exg, "right", return_attention=True)]
        ua = [t.clone() for t in model(u8, exg, return_attention=True)]
    finally:
        model.input_size = None
    xf = model.preprocess(u8, 224)
    torch.manual_seed(11)
    b = model.forward_with_guidance(xf, exg, "right", return_attention=True)
    ub = model(xf, exg, return_attention=True)
    for p, q in list(zip(a, b)) + list(zip(ua, ub)):
        assert (p - q).abs().max().item() <= 2e-3 * max(1.0, q.abs().max().item())
    assert torch.equal(a[2].argmax(-1), b[2].argmax(-1))
    with pytest.raises(ValueError):
        model.forward_with_guidance(u8, exg, "right")          # not square and no input_size


def test_layernorm_folded_backbone_and_guided_forward_match_the_oracle(cuda_device, sd, cases, monkeypatch):
    """CA_LN_FOLD=1: norm1 / norm2 folded into the GEMMs either side of them (no LayerNorm pass inside the encoder) — the
    same oracle tolerances as the default path."""
    from cognitive_aim_depth_estimation_b200.model import create_model
    monkeypatch.setenv("CA_LN_FOLD", "1")
    m = create_model(CFG, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    assert m._ln_folded()
    for (S, B), (x, ex, ref_tok) in cases.items():
        tok = m.backbone_tokens(x.cuda()).cpu()
        rel = ((tok - ref_tok).norm() / ref_tok.norm()).item()
        assert torch.isfinite(tok).all() and rel < TOKEN_REL_FRO, rel
        torch.manual_seed(11)
        ref = orc.forward_with_guidance(sd, None, ex, "left", tokens=ref_tok, update_history=False)
        torch.manual_seed(11)
        depth, conf, heat = m.forward_with_guidance(x.cuda(), _cuda_exif(ex, "cuda"), "left", return_attention=True)
        assert ((depth.cpu() - ref["depth"]).abs() / ref["depth"].abs()).max().item() <= DEPTH_ABS_REL
        assert (conf.cpu() - ref["confidence"]).abs().max().item() <= CONF_MAX_ABS
        assert (heat.cpu() - ref["heatmap"]).abs().max().item() <= HEAT_MAX_ABS
        assert torch.equal(heat.cpu().argmax(-1), ref["heatmap"].argmax(-1))


def test_layernorm_folding_with_a_fused_lora_adapter_on_the_attention_output(cuda_device, monkeypatch):
    """CA_LN_FOLD=1 with lora_mode: fused on attention_output: the residual epilogue that feeds norm2 runs with K = 832."""
    from cognitive_aim_depth_estimation_b200.model import create_model
    monkeypatch.setenv("CA_LN_FOLD", "1")
    sd = _lora_sd(0.5)
    merged = dict(sd)
    for i in range(12):
        k = f"backbone.encoder.layer.{i}.attention.output.dense.weight"
        merged[k] = sd[k] + sd[f"lora_layers.{i}.lora_B"] @ sd[f"lora_layers.{i}.lora_A"]
    x = orc.synthetic_images(2, 224)
    want = orc.dinov2_tokens(merged, x)
    m = create_model(dict(CFG, use_lora=True, lora_merge_target="attention_output", lora_mode="fused"),
                     {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    assert m._ln_folded()
    tok = m.backbone_tokens(x.cuda()).cpu()
    assert ((tok - want).norm() / want.norm()).item() < TOKEN_REL_FRO
    # fused adapters on q/k/v need the normalised rows: folding steps aside
    m2 = create_model(dict(CFG, use_lora=True, lora_merge_target="query", lora_mode="fused"), {"num_cameras": 71},
                      device=cuda_device)
    assert not m2._ln_folded()
