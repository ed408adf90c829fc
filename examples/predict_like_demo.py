"""The reference's `CognitiveAimInference.predict` flow (reference demo.py:300-405) on the B200 path.

    python examples/predict_like_demo.py IMAGE.jpg [--instruction center] [--image-size 224] [--overlay out.npy]

What demo.py does per image and what runs here instead:
    Image.open(path).convert('RGB')                    -> nvJPEG decode on the GPU              (model.preprocess_jpeg)
    Resize((S, S)) / ToTensor / Normalize              -> exact-Pillow resample + normalise     (same call)
    default EXIF 50 mm, f/2.8, ISO 100, camera 0       -> the same defaults                    (demo.py:271-277)
    delattr(model, '_last_attention_weights')          -> the same                              (demo.py:334-335)
    model.forward_with_guidance(x, exif, instruction)  -> the same call                         (demo.py:344-346)
    numpy / scipy heat-map overlay                     -> model.focus_map((h, w)) on the GPU    (demo.py:530-563)
Weights: random init (seed 0, the reference's own construction order) unless --checkpoint names a reference state_dict.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cognitive_aim_depth_estimation_b200.model import create_model  # noqa: E402

INSTRUCTIONS = ["center", "left", "right", "top", "bottom", "top-left", "top-right", "bottom-left", "bottom-right"]
CONFIG = {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"],  # demo.py:46-52
          "model": {"exif_config": {"num_cameras": 71}}}


class Predictor:
    """Counterpart of demo.py's CognitiveAimInference: one model on one GPU, `predict(jpeg bytes, instruction)`."""

    def __init__(self, state_dict=None, device="cuda:0", image_size=224):
        self.device, self.image_size = torch.device(device), image_size
        self.model = create_model(CONFIG, {"num_cameras": 71}, device=self.device)  # demo.py:56-68
        if state_dict is not None:
            self.model.load_state_dict(state_dict, strict=False)                    # demo.py:89-150
        self.model.eval()

    def default_exif(self, n=1):
        """demo.py:271-277: no EXIF in the file -> 50 mm, f/2.8, ISO 100, camera index 0."""
        f = dict(device=self.device, dtype=torch.float32)
        return {"focal_length": torch.full((n,), 50.0, **f), "aperture": torch.full((n,), 2.8, **f),
                "iso": torch.full((n,), 100.0, **f), "camera_idx": torch.zeros(n, device=self.device, dtype=torch.long)}

    @torch.no_grad()
    def predict(self, jpeg_bytes: bytes, instruction=None, exif=None, overlay_size=None):
        """-> (depth, confidence, metadata) like demo.py:300-405; metadata carries the attention arg-max cell and, when
        `overlay_size` = (h, w) is given, the overlay heat map demo.py would paint."""
        x = self.model.preprocess_jpeg([jpeg_bytes], self.image_size)
        exif = exif if exif is not None else self.default_exif(1)
        if hasattr(self.model, "_last_attention_weights"):
            delattr(self.model, "_last_attention_weights")
        if instruction is not None:
            depth, conf = self.model.forward_with_guidance(x, exif, instruction)
        else:
            depth, conf = self.model(x, exif)
        att = self.model.get_attention_weights()
        g = self.image_size // 14
        cell = int(att[0].argmax())
        meta = {"processed_size": (self.image_size, self.image_size), "instruction": instruction,
                "attention_cell": (cell // g, cell % g),
                "model_status": {"ambient": self.model.use_ambient, "focal": self.model.use_focal, "exif": self.model.use_exif}}
        if overlay_size is not None:
            meta["overlay"] = self.model.focus_map(overlay_size)[0]
        return float(depth.squeeze()), float(conf.squeeze()), meta


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("image")
    ap.add_argument("--instruction", default=None, help="one of %s, or 'all'" % ", ".join(INSTRUCTIONS))
    ap.add_argument("--image-size", type=int, default=224)  # configs/experiment_B.yaml:87
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--overlay", default=None, help="write the overlay heat map of the last instruction as .npy")
    args = ap.parse_args()
    # without a checkpoint: the reference's own random init (demo.py:148-150 "continuing with randomly initialized
    # weights") — create_model draws it from the global generator exactly like the reference's constructor
    sd = torch.load(args.checkpoint, map_location="cpu") if args.checkpoint else None
    torch.manual_seed(0)
    p = Predictor(sd, image_size=args.image_size)
    data = open(args.image, "rb").read()
    todo = INSTRUCTIONS if args.instruction == "all" else [args.instruction]
    for ins in todo:
        depth, conf, meta = p.predict(data, ins, overlay_size=(480, 640) if args.overlay else None)
        print(f"{str(ins):13s} depth {depth:.4f}  confidence {conf:.4f}  attention cell {meta['attention_cell']}")
    if args.overlay:
        import numpy as np
        np.save(args.overlay, meta["overlay"].cpu().numpy())


if __name__ == "__main__":
    main()
