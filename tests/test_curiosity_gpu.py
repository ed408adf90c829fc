"""CuriosityModule on the GPU: rewards, the exploration ring buffer and its pointer
— observable `state_dict` state of the reference — and the curiosity-guided attention configuration, against the oracle
and against fixtures recorded from the unmodified reference (tests/golden/curiosity*.npz, oracle/make_golden.py §6-7)."""
import os

import numpy as np
import pytest
import torch

from oracle import cogaim_oracle as orc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
CFG_GUIDED = dict(CFG, curiosity_guided_attention={"enabled": True})  # top level: where the reference looks (:854)
REWARD_RTOL = 1e-2  # bf16 backbone -> CLS token -> fp32 curiosity MLPs
DEPTH_ABS_REL, HEAT_MAX_ABS = 1e-2, 1e-2


def _exif(ex):
    return {k: v.cuda() for k, v in ex.items()}


def _model(cuda_device, cfg, sd):
    from cognitive_aim_depth_estimation_b200.model import create_model
    m = create_model(cfg, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    return m


def test_curiosity_kernel_matches_oracle(cuda_device):
    """ca_curiosity on given fp32 CLS rows and given draws vs the oracle restatement: fp32 both sides."""
    from cognitive_aim_depth_estimation_b200 import ops
    sd = orc.build_state_dict(0)
    m = _model(cuda_device, CFG, sd)
    pk = m._pack()
    B, T = 37, 3
    g = torch.Generator().manual_seed(3)
    tokens = torch.randn(B, T, 768, generator=g) * 1.5
    torch.manual_seed(5)
    eps, noise = torch.randn(B, 192), torch.randn(B, 768)
    s2 = dict(sd)
    s2["curiosity_module.exploration_history"] = torch.zeros(1000)
    s2["curiosity_module.history_pointer"] = torch.tensor(990)
    torch.manual_seed(5)
    want = orc.curiosity_module(s2, tokens[:, 0])
    hist = torch.zeros(1000, device="cuda")
    ptr = torch.tensor(990, device="cuda")
    raw, rew = torch.empty(B, device="cuda"), torch.empty(B, device="cuda")
    ops.curiosity(pk["curiosity"], tokens=tokens.cuda(), tokens_per_img=T, eps=eps.cuda(), noise=noise.cuda(),
                  reward_raw=raw, reward=rew, history=hist, history_pointer=ptr, B=B)
    np.testing.assert_allclose(rew.cpu().numpy(), want.numpy(), rtol=2e-5)
    np.testing.assert_allclose(hist.cpu().numpy(), s2["curiosity_module.exploration_history"].numpy(), rtol=2e-5)
    assert int(ptr) == int(s2["curiosity_module.history_pointer"]) == (990 + B) % 1000
    assert float(want.std()) > 1e-3  # the rewards are not a constant


def _replay_sequence(m, gold, check_outputs):
    """The call sequence of oracle/make_golden.py `run_sequence` on the GPU model."""
    x, ex = orc.synthetic_images(2, 224).cuda(), _exif(orc.synthetic_exif(2))
    cm = m.curiosity_module
    cm.exploration_history.zero_()
    cm.history_pointer.zero_()

    def check(step, n_runs):
        assert int(cm.history_pointer) == int(gold[f"seq{step}_pointer"]), step
        np.testing.assert_allclose(cm.exploration_history[:32].cpu().numpy(), gold[f"seq{step}_history"], rtol=REWARD_RTOL)
        np.testing.assert_allclose(cm.exploration_history[-4:].cpu().numpy(), gold[f"seq{step}_history_tail"],
                                   rtol=REWARD_RTOL)
        assert gold[f"seq{step}_rewards"].shape[0] == n_runs

    def close(depth, heat, step):
        if not check_outputs:
            return
        gd, gh = gold[f"seq{step}_depth"], gold[f"seq{step}_heat"]
        assert (np.abs(depth.cpu().numpy() - gd) / np.abs(gd)).max() <= DEPTH_ABS_REL
        assert np.abs(heat.cpu().numpy() - gh).max() <= HEAT_MAX_ABS
        assert (heat.cpu().numpy().argmax(-1) == gh.argmax(-1)).all()

    if hasattr(m, "_last_attention_weights"):
        delattr(m, "_last_attention_weights")
    torch.manual_seed(11)
    d, c, h = m.forward_with_guidance(x, ex, "center", return_attention=True)
    check(1, 1)
    close(d, h, 1)
    np.testing.assert_allclose(m._last_curiosity.cpu().numpy(), gold["seq1_rewards"][0], rtol=REWARD_RTOL)
    torch.manual_seed(11)
    d, c, h = m(x, ex, return_attention=True)          # attention stored: get_features_aligned + return_attention runs
    check(2, 2)
    close(d, h, 2)
    delattr(m, "_last_attention_weights")
    torch.manual_seed(11)
    d, c, h = m(x, None, return_attention=True)        # nothing stored: three runs
    check(3, 3)
    close(d, h, 3)
    torch.manual_seed(11)
    m(x, ex)
    check(4, 1)
    torch.manual_seed(11)
    d, c, h = m.forward_with_guidance(x, None, "left", return_attention=True)   # guided attempt + fallback: three runs
    check(5, 3)
    close(d, h, 5)
    gl = gold["seq5_last_attention"]
    assert np.abs(m.get_attention_weights().cpu().numpy() - gl).max() <= HEAT_MAX_ABS
    assert (m.get_attention_weights().cpu().numpy().argmax(-1) == gl.argmax(-1)).all()
    cm.history_pointer.fill_(999)                       # ring-buffer wrap-around
    torch.manual_seed(11)
    m.forward_with_guidance(x, ex, "top", return_attention=True)
    check(6, 1)


@pytest.mark.parametrize("graphs", [True, False])
def test_exploration_history_follows_the_reference(cuda_device, graphs):
    """Rewards, ring-buffer contents and pointer after each call of a mixed guided / un-guided sequence equal what the
    unmodified reference recorded (1, 2, 3, 1, 3 CuriosityModule runs per call, then a wrap-around) — eagerly and
    through CUDA-graph replay (the buffers are module state the captured kernels write in place)."""
    gold = np.load(os.path.join(GOLD, "curiosity.npz"))
    m = _model(cuda_device, CFG, orc.build_state_dict(0))
    m.use_cuda_graphs = graphs
    _replay_sequence(m, gold, check_outputs=False)
    _replay_sequence(m, gold, check_outputs=False)  # second round: every graph is now a replay
    sd = m.state_dict()
    assert int(sd["curiosity_module.history_pointer"]) == 1 and sd["curiosity_module.history_pointer"].dtype == torch.int64


def test_curiosity_guided_configuration(cuda_device):
    """`curiosity_guided_attention: {enabled: true}` at the top level of the config: amplifier + modulators scale the
    focal attention before its clamp / renormalisation (src/model.py:264-276, :406-417)."""
    gold = np.load(os.path.join(GOLD, "curiosity_guided.npz"))
    sd = orc.build_state_dict(0, curiosity_guided=True)
    m = _model(cuda_device, CFG_GUIDED, sd)
    assert m.cfg.curiosity_guided and len(m.state_dict()) == 335
    _replay_sequence(m, gold, check_outputs=True)
    x, ex = orc.synthetic_images(2, 224).cuda(), _exif(orc.synthetic_exif(2))
    for ins in ("top-left", "right"):
        torch.manual_seed(11)
        d, c, h = m.forward_with_guidance(x, ex, ins, return_attention=True)
        gd, gh = gold[f"guided_{ins}_depth"], gold[f"guided_{ins}_heat"]
        assert (np.abs(d.cpu().numpy() - gd) / np.abs(gd)).max() <= DEPTH_ABS_REL
        assert np.abs(h.cpu().numpy() - gh).max() <= HEAT_MAX_ABS
        assert (h.cpu().numpy().argmax(-1) == gh.argmax(-1)).all()


def test_curiosity_modulation_kernel_matches_oracle(cuda_device):
    """ca_curiosity_modulation + the modulated ca_focal_finalize vs the oracle's focal stream on oracle tokens, with an
    adaptive weight and modulator biases pushed away from their init so that the scale factor is not ~1."""
    sd = orc.build_state_dict(0, curiosity_guided=True)
    g = torch.Generator().manual_seed(9)
    for i in range(3):
        p = f"focal_stream.focal_streams.{i}."
        sd[p + "adaptive_weight"] = torch.tensor(0.3 + 0.2 * i)
        sd[p + "curiosity_modulator.2.bias"] = torch.randn(8, generator=g)
    sd["focal_stream.curiosity_amplifier.2.bias"] = torch.tensor([0.5, -0.3, 0.1])
    m = _model(cuda_device, CFG_GUIDED, sd)
    tokens = orc.dinov2_tokens(sd, orc.synthetic_images(2, 224))
    score = torch.tensor([0.37, 1.9])
    _, want = orc.iterative_focal_stream(sd, tokens[:, 1:], need_features=False, curiosity_score=score)
    from cognitive_aim_depth_estimation_b200 import ops
    pk = m._pack()
    ws = m._workspace(2, 224)
    ws["tokens"].copy_(tokens.cuda())
    cw = torch.empty(3, 2, device="cuda")
    ops.curiosity_modulation(pk["curiosity_mod"], score.cuda(), 1.0, 0.0, cw, 2, 3, 32)
    # oracle-side modulation weights
    iw = torch.softmax(orc._mlp(sd, score.unsqueeze(-1), ["focal_stream.curiosity_amplifier.0",
                                                            "focal_stream.curiosity_amplifier.2"]), dim=-1)
    for i in range(3):
        q = f"focal_stream.focal_streams.{i}.curiosity_modulator."
        ref = torch.sigmoid(orc._mlp(sd, (score * iw[:, i]).unsqueeze(-1), [q + "0", q + "2"])).mean(-1)
        np.testing.assert_allclose(cw[i].cpu().numpy(), ref.numpy(), rtol=1e-5)
    att = m._focal_iterations(ws, 2, 16, want_features=False, cur_weight=cw).cpu()
    rel = (att - want).abs() / want
    assert rel.max().item() < 5e-2 and rel.mean().item() < 2e-3
    clamped = torch.empty(3, 2, device="cuda")
    ops.curiosity_modulation(pk["curiosity_mod"], score.cuda(), 0.5, 1.0, clamped, 2, 3, 32)
    both = torch.empty(3, 2, device="cuda")
    ops.curiosity_modulation(pk["curiosity_mod"], score.clamp(0.5, 1.0).cuda(), 1.0, 0.0, both, 2, 3, 32)
    assert torch.equal(clamped, both)
