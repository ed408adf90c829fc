"""Pin the CPU oracle (oracle/cogaim_oracle.py) against outputs of the UNMODIFIED reference.

The fixtures in tests/golden/ were produced by oracle/make_golden.py, which imports /root/reference in the build
container.  These tests run on CPU (`-m "not gpu"`) and never touch /root/reference.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cogaim_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def sd():
    return orc.build_state_dict(0)


@pytest.fixture(scope="module")
def tokens224(sd):
    return orc.dinov2_tokens(sd, orc.synthetic_images(2, 224))


def test_state_dict_matches_reference_init(sd):
    """Same seed + same construction order => the reference's 319 tensors, bit for bit (digest compare)."""
    gold = json.load(open(os.path.join(GOLD, "state_dict_seed0.json")))
    assert list(sd.keys()) == gold["names"]
    assert len(sd) == 319
    for k, v in sd.items():
        assert list(v.shape) == gold["shapes"][k], k
        s, a = gold["digest"][k]
        assert float(v.double().sum()) == s and float(v.double().abs().sum()) == a, k


def test_backbone_restated_vs_hf(sd, tokens224):
    gold = np.load(os.path.join(GOLD, "backbone.npz"))
    t = tokens224
    np.testing.assert_allclose(t[:, :8, :32].numpy(), gold["S224_B2_tokens_head"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(t[:, -4:, -32:].numpy(), gold["S224_B2_tokens_tail"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(t.norm(dim=-1).numpy(), gold["S224_B2_token_norms"], rtol=1e-5)


@pytest.mark.parametrize("instruction", orc.INSTRUCTIONS + ["TopLeft", "CENTER", "nonsense"])
def test_guided_224(sd, tokens224, instruction):
    gold = np.load(os.path.join(GOLD, "guided.npz"))
    torch.manual_seed(11)
    out = orc.forward_with_guidance(sd, None, orc.synthetic_exif(2), instruction, tokens=tokens224,
                                    update_history=False)
    key = f"S224_B2_{instruction}"
    np.testing.assert_allclose(out["depth"].numpy(), gold[key + "_depth"], rtol=2e-5)
    np.testing.assert_allclose(out["confidence"].numpy(), gold[key + "_conf"], rtol=1e-6)
    np.testing.assert_allclose(out["heatmap"].numpy(), gold[key + "_heat"], rtol=2e-3, atol=1e-9)
    assert (out["heatmap"].argmax(-1).numpy() == gold[key + "_heat"].argmax(-1)).all()


def test_guided_tensor_guidance_is_resized(sd, tokens224):
    gold = np.load(os.path.join(GOLD, "guided.npz"))
    torch.manual_seed(11)
    out = orc.forward_with_guidance(sd, None, orc.synthetic_exif(2), torch.linspace(0.5, 4.0, 196), tokens=tokens224,
                                    update_history=False)
    np.testing.assert_allclose(out["heatmap"].numpy(), gold["S224_B2_tensor196_heat"], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(out["depth"].numpy(), gold["S224_B2_tensor196_depth"], rtol=2e-5)


def test_guided_518_center_and_corner(sd):
    gold = np.load(os.path.join(GOLD, "guided.npz"))
    tokens = orc.dinov2_tokens(sd, orc.synthetic_images(1, 518))
    bg = np.load(os.path.join(GOLD, "backbone.npz"))
    np.testing.assert_allclose(tokens.norm(dim=-1).numpy(), bg["S518_B1_token_norms"], rtol=1e-5)
    for ins in ("center", "top-left", "bottom-right"):
        torch.manual_seed(11)
        out = orc.forward_with_guidance(sd, None, orc.synthetic_exif(1), ins, tokens=tokens, update_history=False)
        key = f"S518_B1_{ins}"
        np.testing.assert_allclose(out["depth"].numpy(), gold[key + "_depth"], rtol=2e-5)
        np.testing.assert_allclose(out["heatmap"].numpy(), gold[key + "_heat"], rtol=2e-3, atol=1e-9)
        assert (out["heatmap"].argmax(-1).numpy() == gold[key + "_heat"].argmax(-1)).all()


@pytest.mark.parametrize("tag", ["exif", "noexif"])
def test_unguided_224(sd, tokens224, tag):
    gold = np.load(os.path.join(GOLD, "unguided.npz"))
    torch.manual_seed(11)
    ex = orc.synthetic_exif(2) if tag == "exif" else None
    out = orc.forward_unguided(sd, None, ex, tokens=tokens224, update_history=False)
    key = f"S224_B2_{tag}"
    np.testing.assert_allclose(out["depth"].numpy(), gold[key + "_depth"], rtol=2e-5)
    np.testing.assert_allclose(out["confidence"].numpy(), gold[key + "_conf"], rtol=1e-6)
    np.testing.assert_allclose(out["heatmap"].numpy(), gold[key + "_heat"], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(out["fusion_features"].numpy(), gold[key + "_fusion"], rtol=1e-4, atol=1e-6)


def test_confidence_degenerate_constant():
    """SURVEY.md §4 degeneracy guard: at seed-0 init every confidence is sigmoid(2.0)."""
    gold = np.load(os.path.join(GOLD, "guided.npz"))
    assert np.allclose(gold["S224_B2_center_conf"], 0.8807970285, atol=1e-7)
