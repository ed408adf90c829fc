"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference) on CPU fp32.

Runs only in the build container (the reference does not travel to the GPU box).  The only patch is the
network access at reference src/model.py:814 (`Dinov2Model.from_pretrained`) -> random-init
`Dinov2Model(Dinov2Config(image_size=518, patch_size=14))`.  Seed protocol: weights
`torch.manual_seed(0)`; images seed 1234; EXIF seed 1236; `torch.manual_seed(11)` before every call.

    python oracle/make_golden.py                    # writes tests/golden/*.npz / *.json
    python oracle/make_golden.py --only-curiosity   # only curiosity.npz / curiosity_guided.npz (sections 6-7)
    python oracle/make_golden.py --only-lora        # only lora.npz / state_dict_seed0_lora.json (section 8)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import cogaim_oracle as orc  # noqa: E402  (inputs + digest helper only)

WEIGHT_SEED, CALL_SEED = 0, 11


def load_reference():
    from transformers import Dinov2Config, Dinov2Model
    Dinov2Model.from_pretrained = staticmethod(
        lambda name, *a, **k: Dinov2Model(Dinov2Config(image_size=518, patch_size=14)))
    sys.path.insert(0, REF)
    import src.model as ref  # noqa
    return ref


def make_model(ref, cfg_path, extra=None):
    cfg = yaml.safe_load(open(cfg_path))
    cfg.update(extra or {})
    cfg.setdefault("cognitive_modules", ["ambient_stream", "iterative_focal_stream", "exif_prior_database"])  # demo.py:46-52
    torch.manual_seed(WEIGHT_SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.create_model(cfg, {"num_cameras": 71}).eval()
    return model


def effective_attrs(model):
    fs = model.focal_stream
    return {
        "use_lora": bool(model.use_lora), "use_ambient": bool(model.use_ambient), "use_focal": bool(model.use_focal),
        "use_exif": bool(model.use_exif), "use_iterative": bool(model.use_iterative),
        "feature_dim": int(model.feature_dim), "fusion_dim": int(model.fusion_dim),
        "num_iterations": int(fs.num_iterations), "focus_strength": float(fs.focus_strength),
        "curiosity_guided": bool(fs.curiosity_guided),
        "n_state_tensors": len(model.state_dict()),
        "n_params": int(sum(p.numel() for p in model.parameters())),
    }


def curiosity_sections(ref):
    """6. CuriosityModule side effects (rewards, exploration ring buffer, pointer) over a sequence of calls, and
    7. the curiosity-guided configuration (top-level `curiosity_guided_attention.enabled`, src/model.py:854)."""
    cfg_path = os.path.join(REF, "configs", "experiment_B.yaml")
    S, B = 224, 2
    x, ex = orc.synthetic_images(B, S), orc.synthetic_exif(B)

    def run_sequence(model, out, tag):
        rewards = []
        hook = model.curiosity_module.register_forward_hook(lambda m, i, o: rewards.append(o[0].detach().clone()))
        cm = model.curiosity_module

        def snap(step):
            out[f"{tag}{step}_history"] = cm.exploration_history[:32].numpy().copy()
            out[f"{tag}{step}_history_tail"] = cm.exploration_history[-4:].numpy().copy()
            out[f"{tag}{step}_pointer"] = np.asarray(int(cm.history_pointer))
            out[f"{tag}{step}_rewards"] = torch.stack(rewards).numpy() if rewards else np.zeros((0, B), np.float32)
            rewards.clear()

        def call(fn, *a, **k):
            torch.manual_seed(CALL_SEED)
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                return fn(*a, **k)

        if hasattr(model, "_last_attention_weights"):
            delattr(model, "_last_attention_weights")
        d, c, h = call(model.forward_with_guidance, x, ex, "center", return_attention=True)   # 1 curiosity run
        out[f"{tag}1_depth"], out[f"{tag}1_heat"] = d.numpy(), h.numpy()
        snap(1)
        d, c, h = call(model, x, ex, return_attention=True)      # _last_attention_weights present: 2 runs
        out[f"{tag}2_depth"], out[f"{tag}2_heat"], out[f"{tag}2_fusion"] = d.numpy(), h.numpy(), model.fusion_features.numpy()
        snap(2)
        delattr(model, "_last_attention_weights")
        d, c, h = call(model, x, None, return_attention=True)    # absent: 3 runs
        out[f"{tag}3_depth"], out[f"{tag}3_heat"] = d.numpy(), h.numpy()
        snap(3)
        call(model, x, ex)                                       # present, no attention: 1 run
        snap(4)
        d, c, h = call(model.forward_with_guidance, x, None, "left", return_attention=True)   # guided attempt + fallback
        out[f"{tag}5_depth"], out[f"{tag}5_heat"] = d.numpy(), h.numpy()
        out[f"{tag}5_last_attention"] = model._last_attention_weights.numpy()
        snap(5)
        cm.history_pointer.fill_(999)                            # wrap-around of the ring buffer
        call(model.forward_with_guidance, x, ex, "top", return_attention=True)
        snap(6)
        hook.remove()

    out = {}
    run_sequence(make_model(ref, cfg_path), out, "seq")
    np.savez_compressed(os.path.join(OUT, "curiosity.npz"), **out)
    for k in sorted(out):
        if k.endswith("_pointer") or k.endswith("_rewards"):
            print(k, out[k].ravel())

    out = {}
    model = make_model(ref, cfg_path, {"curiosity_guided_attention": {"enabled": True}})
    assert model.focal_stream.curiosity_guided
    sd = model.state_dict()
    json.dump({"seed": WEIGHT_SEED, "names": list(sd.keys()), "shapes": {k: list(v.shape) for k, v in sd.items()},
               "digest": orc.state_dict_digest(sd)}, open(os.path.join(OUT, "state_dict_seed0_curiosity_guided.json"), "w"))
    run_sequence(model, out, "seq")
    for ins in ("top-left", "right"):
        if hasattr(model, "_last_attention_weights"):
            delattr(model, "_last_attention_weights")
        torch.manual_seed(CALL_SEED)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            d, c, h = model.forward_with_guidance(x, ex, ins, return_attention=True)
        out[f"guided_{ins}_depth"], out[f"guided_{ins}_conf"], out[f"guided_{ins}_heat"] = d.numpy(), c.numpy(), h.numpy()
    np.savez_compressed(os.path.join(OUT, "curiosity_guided.npz"), **out)
    print("curiosity-guided: state tensors", len(sd))


def lora_section(ref):
    """8. top-level `use_lora: true` (src/model.py:822-831): 24 LoRA tensors in the state_dict that no forward reads —
    recorded with lora_B made non-zero, so that "the reference ignores its adapters" is a pinned fact."""
    model = make_model(ref, os.path.join(REF, "configs", "experiment_B.yaml"), {"use_lora": True})
    assert model.use_lora
    sd = model.state_dict()
    json.dump({"seed": WEIGHT_SEED, "names": list(sd.keys()), "shapes": {k: list(v.shape) for k, v in sd.items()},
               "digest": orc.state_dict_digest(sd)}, open(os.path.join(OUT, "state_dict_seed0_lora.json"), "w"))
    g = torch.Generator().manual_seed(77)
    with torch.no_grad():
        for lo in model.lora_layers:
            lo.lora_B.copy_(torch.randn(lo.lora_B.shape, generator=g) * 0.05)
    x, ex = orc.synthetic_images(2, 224), orc.synthetic_exif(2)
    torch.manual_seed(CALL_SEED)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        d, c, h = model.forward_with_guidance(x, ex, "center", return_attention=True)
    np.savez_compressed(os.path.join(OUT, "lora.npz"), depth=d.numpy(), conf=c.numpy(), heat=h.numpy())
    print("lora: state tensors", len(sd), "depth", d.ravel())


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_num_threads(os.cpu_count())
    if "--only-curiosity" in sys.argv:
        curiosity_sections(ref)
        return
    if "--only-lora" in sys.argv:
        lora_section(ref)
        return

    # 1. effective config of every shipped YAML (quirk 1) ------------------------------------------------
    attrs = {}
    cfgs = [os.path.join(REF, "configs", "experiment_B.yaml")] + sorted(
        os.path.join(REF, "eval_configs", f) for f in os.listdir(os.path.join(REF, "eval_configs")))
    model = None
    for c in cfgs:
        m = make_model(ref, c)
        attrs[os.path.relpath(c, REF)] = effective_attrs(m)
        if model is None:
            model = m
    json.dump(attrs, open(os.path.join(OUT, "effective_config.json"), "w"), indent=1, sort_keys=True)

    # 2. weights digest ----------------------------------------------------------------------------------
    sd = model.state_dict()
    dig = orc.state_dict_digest(sd)
    json.dump({"seed": WEIGHT_SEED, "names": list(sd.keys()), "shapes": {k: list(v.shape) for k, v in sd.items()},
               "digest": dig}, open(os.path.join(OUT, "state_dict_seed0.json"), "w"))

    # 3. guided forward, all 9 instructions ----------------------------------------------------------------
    def guided(S, B, instr, call_seed=CALL_SEED):
        x = orc.synthetic_images(B, S)
        ex = orc.synthetic_exif(B)
        if hasattr(model, "_last_attention_weights"):
            delattr(model, "_last_attention_weights")  # demo.py:334-335
        torch.manual_seed(call_seed)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            d, c, h = model.forward_with_guidance(x, ex, instr, return_attention=True)
        return d.numpy(), c.numpy(), h.numpy()

    out = {}
    for S, B in ((224, 2), (518, 1)):
        for ins in orc.INSTRUCTIONS:
            d, c, h = guided(S, B, ins)
            key = f"S{S}_B{B}_{ins}"
            out[key + "_depth"], out[key + "_conf"], out[key + "_heat"] = d, c, h
            print(key, d.ravel(), c.ravel(), h.argmax(-1))
    # alias / case / unknown-string / tensor-guidance semantics at S=224
    for ins in ("TopLeft", "CENTER", "nonsense"):
        d, c, h = guided(224, 2, ins)
        key = f"S224_B2_{ins}"
        out[key + "_depth"], out[key + "_conf"], out[key + "_heat"] = d, c, h
    gvec = torch.linspace(0.5, 4.0, 196)  # 14x14 guidance on a 16x16 grid -> bilinear resize (:1386-1398)
    d, c, h = guided(224, 2, gvec)
    out["S224_B2_tensor196_depth"], out["S224_B2_tensor196_conf"], out["S224_B2_tensor196_heat"] = d, c, h
    np.savez_compressed(os.path.join(OUT, "guided.npz"), **out)

    # 4. un-guided forward (with and without EXIF) -----------------------------------------------------------
    out = {}
    for S, B in ((224, 2),):
        x = orc.synthetic_images(B, S)
        for tag, ex in (("exif", orc.synthetic_exif(B)), ("noexif", None)):
            if hasattr(model, "_last_attention_weights"):
                delattr(model, "_last_attention_weights")
            torch.manual_seed(CALL_SEED)
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                d, c, h = model(x, ex, return_attention=True)
            key = f"S{S}_B{B}_{tag}"
            out[key + "_depth"], out[key + "_conf"], out[key + "_heat"] = d.numpy(), c.numpy(), h.numpy()
            out[key + "_fusion"] = model.fusion_features.numpy()
            print(key, d.ravel(), c.ravel())
    np.savez_compressed(os.path.join(OUT, "unguided.npz"), **out)

    # 5. backbone tokens (HF Dinov2Model inside the reference model) ------------------------------------------
    out = {}
    for S, B in ((224, 2), (518, 1)):
        x = orc.synthetic_images(B, S)
        with torch.no_grad():
            t = model.backbone(x, output_hidden_states=True).last_hidden_state
        out[f"S{S}_B{B}_tokens_head"] = t[:, :8, :32].numpy()
        out[f"S{S}_B{B}_tokens_tail"] = t[:, -4:, -32:].numpy()
        out[f"S{S}_B{B}_token_norms"] = t.norm(dim=-1).numpy()
    np.savez_compressed(os.path.join(OUT, "backbone.npz"), **out)
    curiosity_sections(ref)
    lora_section(ref)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
