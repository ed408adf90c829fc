"""Bandwidth-bound row kernels (csrc/rowops.cu) against their torch fp32 definitions."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _unfold_ref(images):
    """[B,3,S,S] -> [B*g*g, 588] with column k = c*196 + ky*14 + kx (Conv2d weight.flatten(1) order)."""
    B, C, S, _ = images.shape
    g = S // 14
    x = images[:, :, :g * 14, :g * 14].reshape(B, C, g, 14, g, 14).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B * g * g, C * 196)


@pytest.mark.parametrize("B,S", [(2, 224), (1, 518), (3, 70)])
def test_patchify_f32(cuda_device, B, S):
    from cognitive_aim_depth_estimation_b200 import ops
    img = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    g = S // 14
    patches = torch.full((B * g * g, ops.PATCH_ROW_STRIDE), 7.0, device=cuda_device, dtype=torch.bfloat16)
    ops.patchify_f32(img, patches)
    assert torch.equal(patches[:, :588], _unfold_ref(img).bfloat16())
    assert (patches[:, 588:] == 0).all()


def test_patchify_equals_conv(cuda_device):
    """patch rows x flattened conv weight == the reference's Conv2d(3,768,14,14) (HF modeling_dinov2.py:139-148)."""
    from cognitive_aim_depth_estimation_b200 import ops
    img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    w = torch.randn(768, 3, 14, 14, generator=torch.Generator().manual_seed(3)).to(cuda_device) * 0.02
    patches = torch.zeros((2 * 256, ops.PATCH_ROW_STRIDE), device=cuda_device, dtype=torch.bfloat16)
    ops.patchify_f32(img, patches)
    got = patches[:, :588].float() @ w.flatten(1).t()
    ref = F.conv2d(img.bfloat16().float(), w, stride=14).flatten(2).transpose(1, 2).reshape(-1, 768)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)


def test_preprocess_u8(cuda_device):
    from cognitive_aim_depth_estimation_b200 import ops
    B, S = 2, 224
    u8 = torch.randint(0, 256, (B, S, S, 3), generator=torch.Generator().manual_seed(1235), dtype=torch.uint8)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    ref_img = (u8.permute(0, 3, 1, 2).float() / 255.0 - mean) / std  # ToTensor + Normalize (demo.py:162-166)
    patches = torch.zeros((B * 256, ops.PATCH_ROW_STRIDE), device=cuda_device, dtype=torch.bfloat16)
    ops.preprocess_u8(u8.to(cuda_device), patches)
    ref = _unfold_ref(ref_img).to(cuda_device)
    # (x/255 - m) * (1/s) vs (x/255 - m) / s differ by 1 ulp fp32 at most -> identical after bf16 rounding except ties
    diff = (patches[:, :588].float() - ref.bfloat16().float()).abs()
    assert (diff <= ref.abs() * 2 ** -7).all()
    assert (diff > 0).float().mean() < 1e-3


@pytest.mark.parametrize("rows", [1, 37, 1370 * 3])
def test_layernorm(cuda_device, rows):
    from cognitive_aim_depth_estimation_b200 import ops
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, 768, generator=g) * 3 + 1.5).to(cuda_device)
    gamma = (torch.randn(768, generator=g) * 0.2 + 1).to(cuda_device)
    beta = (torch.randn(768, generator=g) * 0.1).to(cuda_device)
    ref = F.layer_norm(x, (768,), gamma, beta, 1e-6)
    o32 = torch.empty_like(x)
    ops.layernorm(x, gamma, beta, o32)
    assert torch.allclose(o32, ref, rtol=1e-5, atol=2e-6)
    o16 = torch.empty((rows, 768), device=cuda_device, dtype=torch.bfloat16)
    ops.layernorm(x, gamma, beta, o16)
    assert torch.equal(o16, o32.bfloat16())


def test_cls_rows_and_focal_input(cuda_device):
    from cognitive_aim_depth_estimation_b200 import ops
    B, N, D = 3, 256, 768
    g = torch.Generator().manual_seed(5)
    tokens = torch.randn(B, N + 1, D, generator=g).to(cuda_device)
    pe = torch.randn(N, D, generator=g).to(cuda_device)
    rs = (torch.rand(B, N, generator=g) + 1).to(cuda_device)
    xin = torch.empty((B * N, D), device=cuda_device, dtype=torch.bfloat16)
    ops.focal_input(tokens, pe, rs, xin, B, N, D)
    ref = (tokens[:, 1:] * rs.unsqueeze(-1) + pe).reshape(B * N, D)
    assert (xin.float() - ref).abs().max() <= ref.abs().max() * 2 ** -8
    ops.focal_input(tokens, pe, None, xin, B, N, D)
    assert torch.equal(xin, (tokens[:, 1:] + pe).reshape(B * N, D).bfloat16())
    cls = torch.randn(D, generator=g).to(cuda_device)
    pos = torch.randn(N + 1, D, generator=g).to(cuda_device)
    x = torch.zeros(B, N + 1, D, device=cuda_device)
    ops.cls_rows(x, cls, pos, B, N + 1, D)
    assert torch.equal(x[:, 0], (cls + pos[0]).expand(B, D))
    assert (x[:, 1:] == 0).all()


@pytest.mark.parametrize("g,size", [(16, (224, 224)), (37, (518, 518)), (37, (480, 640)), (74, (300, 1036))])
def test_focus_map_matches_numpy_scipy(cuda_device, g, size):
    """Heat-map post-processing for visualisation (demo.py:530-563) on the GPU against the oracle, which runs the
    reference's own numpy / scipy calls."""
    from cognitive_aim_depth_estimation_b200 import ops
    from oracle import cogaim_oracle as orc
    B, N = 3, g * g
    gen = torch.Generator().manual_seed(g)
    heat = torch.softmax(torch.randn(B, N, generator=gen) * 3.0, dim=-1)
    heat[1, :7] = heat[1, 7]  # ties around the order statistics
    want = orc.focus_map(heat, *size)
    norm = torch.empty(B, N, device=cuda_device)
    out = torch.empty(B, *size, device=cuda_device)
    ops.focus_map(heat.to(cuda_device), g, size[0], size[1], norm, out)
    got = out.cpu()
    assert got.shape == want.shape
    assert float(got.min()) >= 0.0 and float(got.max()) <= 1.0
    assert (got - want).abs().max().item() < 2e-6, (got - want).abs().max().item()
    # the normalised grid itself (before the zoom) is exact up to numpy's powf vs x*x*x
    ref_grid = orc.focus_map(heat, g, g)
    assert (norm.cpu().view(B, g, g) - ref_grid).abs().max().item() < 2e-6


@pytest.mark.parametrize("H,W,oh,ow", [(480, 640, 518, 518), (480, 640, 224, 224), (1080, 1920, 518, 518),
                                       (100, 77, 224, 224), (224, 224, 518, 518), (300, 518, 518, 518),
                                       (518, 300, 518, 518), (37, 41, 14, 70), (518, 518, 518, 518)])
def test_resize_matches_pillow(cuda_device, H, W, oh, ow):
    """GPU resize == PIL.Image.resize(..., BILINEAR) bit for bit (demo.py:162-163 via torchvision Resize): down- and
    up-scaling, one-axis-only, identity, ragged windows at the borders."""
    import numpy as np
    from PIL import Image
    from cognitive_aim_depth_estimation_b200 import ops
    rng = np.random.default_rng(H * 7 + W)
    imgs = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    imgs[1, : H // 2] = 255  # saturated region: exercises the 8-bit clamp
    want = np.stack([np.asarray(Image.fromarray(i).resize((ow, oh), Image.BILINEAR)) for i in imgs])
    got = ops.resize_u8(torch.from_numpy(imgs).to(cuda_device), oh, ow).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), int(np.abs(got.astype(int) - want.astype(int)).max())


def test_preprocess_matches_torchvision(cuda_device):
    """model.preprocess == torchvision Compose([Resize((S,S)), ToTensor(), Normalize(imagenet)]) (demo.py:162-166)."""
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from cognitive_aim_depth_estimation_b200.model import create_model
    cfg = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
    m = create_model(cfg, {"num_cameras": 71}, device=cuda_device)
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 256, (2, 480, 640, 3), dtype=np.uint8)
    tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    want = torch.stack([tf(Image.fromarray(i)) for i in imgs])
    got = m.preprocess(torch.from_numpy(imgs), 224).cpu()
    assert torch.equal(got, want)
