"""How far is nvJPEG from Pillow's libjpeg-turbo on the same byte stream?  (sets the tolerance of tests/test_jpeg_gpu.py)"""
import io
import sys
import numpy as np
import torch
from PIL import Image
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops


def synthetic(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([127 + 100 * np.sin(x / 37.0 + seed) * np.cos(y / 53.0), 127 + 90 * np.cos((x + y) / 71.0),
                    255 * (x / w) * (y / h)], -1)
    img[h // 4: h // 2, w // 3: w // 2] = (200, 30, 60)          # a hard-edged block
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


for (h, w) in ((480, 640), (517, 389), (64, 48)):
    for sub, q in ((0, 95), (2, 95), (2, 75), (1, 90)):
        buf = io.BytesIO()
        Image.fromarray(synthetic(h, w, h + sub)).save(buf, format="JPEG", quality=q, subsampling=sub)
        data = buf.getvalue()
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
        got = ops.jpeg_decode(data).cpu().numpy().astype(np.int32)
        d = np.abs(got - ref)
        print(f"{h}x{w} subsampling {sub} q{q}: shape ok {got.shape == ref.shape}  max {d.max()}  mean {d.mean():.4f}  "
              f">1: {(d > 1).mean() * 100:.3f}%  >2: {(d > 2).mean() * 100:.3f}%")
buf = io.BytesIO()
Image.fromarray(synthetic(100, 120, 1)[..., 0]).save(buf, format="JPEG", quality=90)
data = buf.getvalue()
ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB")).astype(np.int32)
got = ops.jpeg_decode(data).cpu().numpy().astype(np.int32)
print("grayscale:", got.shape, ref.shape, np.abs(got - ref).max())
