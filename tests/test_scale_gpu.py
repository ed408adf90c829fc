"""Full-size and scale checks of the B200 path (BASELINE.json configs 2, 4, 5): properties that do not need the CPU
oracle at full size (batch-composition invariance, sharding invariance, graph replay == eager launch, row sums),
plus one oracle parity case at 1036 x 1036 (4x tokens, bicubic position-embedding interpolation, ragged tiles)."""
import pytest
import torch

from oracle import cogaim_oracle as orc

pytestmark = pytest.mark.gpu

CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}


@pytest.fixture(scope="module")
def sd():
    return orc.build_state_dict(0)


@pytest.fixture(scope="module")
def model(cuda_device, sd):
    from cognitive_aim_depth_estimation_b200.model import create_model
    m = create_model(CFG, {"num_cameras": 71}, device=cuda_device)
    m.load_state_dict(sd)
    return m


def _exif(ex, lo=None, hi=None):
    return {k: v[lo:hi].cuda() for k, v in ex.items()}


def _guided(model, x, ex, instruction, replay_batch=None, replay_offset=0):
    torch.manual_seed(11)
    model.rng_replay_batch, model.rng_replay_offset = replay_batch, replay_offset
    try:
        out = model.forward_with_guidance(x, ex, instruction, return_attention=True)
    finally:
        model.rng_replay_batch, model.rng_replay_offset = None, 0
    return [t.clone() for t in out]


def test_full_batch_properties_and_batch_invariance(model):
    """Config 2 (518 x 518, 32 images / GPU): every image's outputs are bit-identical to the ones it gets in a batch of
    2 (no kernel couples images: same tiles, same reduction order), heat-map rows sum to 1, outputs finite."""
    B = 32
    x = orc.synthetic_images(B, 518).cuda()
    ex = orc.synthetic_exif(B)
    depth, conf, heat = _guided(model, x, _exif(ex), "bottom-right")
    assert depth.shape == (B, 1) and conf.shape == (B, 1) and heat.shape == (B, 1369)
    assert torch.isfinite(depth).all() and torch.isfinite(conf).all() and torch.isfinite(heat).all()
    assert (depth > 0).all() and ((conf > 0) & (conf < 1)).all()
    assert torch.allclose(heat.sum(-1), torch.ones(B, device=heat.device), atol=1e-5)
    for lo in (0, 14, 30):
        d2, c2, h2 = _guided(model, x[lo:lo + 2], _exif(ex, lo, lo + 2), "bottom-right", replay_batch=B, replay_offset=lo)
        assert torch.equal(d2, depth[lo:lo + 2]) and torch.equal(c2, conf[lo:lo + 2]) and torch.equal(h2, heat[lo:lo + 2])


def test_sharded_run_equals_unsharded(model):
    """contiguous shards of the global batch, each run as its own call (what each rank of a torchrun
    job does), reproduce the un-sharded result bit for bit, ragged shard included."""
    from cognitive_aim_depth_estimation_b200.sharding import ShardedInference
    B = 7
    x = orc.synthetic_images(B, 224).cuda()
    ex = _exif(orc.synthetic_exif(B))
    want = _guided(model, x, ex, "left")
    for world in (2, 3):
        parts = []
        for rank in range(world):
            torch.manual_seed(11)
            parts.append(ShardedInference(model, rank, world).forward_with_guidance(x, ex, "left", return_attention=True))
        for i in range(3):
            assert torch.equal(torch.cat([p[i] for p in parts]), want[i]), (world, i)


def test_graph_replay_equals_eager(model):
    """The CUDA-graph replay of the forward launches exactly the eager sequence: identical bits."""
    x = orc.synthetic_images(3, 224).cuda()
    ex = _exif(orc.synthetic_exif(3))
    assert model.use_cuda_graphs
    a = _guided(model, x, ex, "top")          # first call: eager + capture
    b = _guided(model, x, ex, "top")          # replay
    c = _guided(model, x, ex, "center")       # replay with another mask through the same fixed buffer
    model.use_cuda_graphs = False
    try:
        d = _guided(model, x, ex, "top")
        e = _guided(model, x, ex, "center")
    finally:
        model.use_cuda_graphs = True
    for i in range(3):
        assert torch.equal(a[i], b[i]) and torch.equal(a[i], d[i]) and torch.equal(c[i], e[i])
    assert not torch.equal(a[2], c[2])
    u1 = [t.clone() for t in model(x, ex, return_attention=True)]
    u2 = [t.clone() for t in model(x, ex, return_attention=True)]
    for i in range(3):
        assert torch.equal(u1[i], u2[i])


def _assert_parity(depth, conf, heat, ref, what):
    """BASELINE.json bar: depth abs-rel <= 1e-2, confidence / heat-map max-abs <= 1e-2, arg-max cell bit-exact.
    Returns the oracle's smallest relative top-1 / top-2 margin over the compared images."""
    depth, conf, heat = depth.cpu(), conf.cpu(), heat.cpu()
    top2 = ref["heatmap"].topk(2, dim=-1).values
    margin = ((top2[:, 0] - top2[:, 1]) / top2[:, 0]).min().item()
    abs_rel = ((depth - ref["depth"]).abs() / ref["depth"].abs()).max().item()
    assert abs_rel <= 1e-2, (what, abs_rel)
    assert (conf - ref["confidence"]).abs().max().item() <= 1e-2, what
    assert (heat - ref["heatmap"]).abs().max().item() <= 1e-2, what
    assert torch.equal(heat.argmax(-1), ref["heatmap"].argmax(-1)), f"{what}: oracle top-1/top-2 margin {margin:.2e}"
    return margin


def test_batch32_images_match_oracle_all_instructions(model, sd):
    """Config 2 AS BENCHMARKED (518 x 518, 32 images in ONE call): images 0, 13, 22 and 31 of the batch against the CPU
    oracle for all 9 instructions.  The oracle runs on those four images only (tokens computed once) but consumes the
    global RNG as the reference would for the whole batch (`rng_rows`), so the per-call random projection is the same."""
    B, rows = 32, [0, 13, 22, 31]
    x = orc.synthetic_images(B, 518)
    ex = orc.synthetic_exif(B)
    tokens = orc.dinov2_tokens(sd, x[rows])
    ex_rows = {k: v[rows] for k, v in ex.items()}
    xg, exg = x.cuda(), _exif(ex)
    for instruction in orc.INSTRUCTIONS:
        torch.manual_seed(11)
        ref = orc.forward_with_guidance(sd, None, ex_rows, instruction, tokens=tokens, update_history=False,
                                        rng_rows=(B, rows))
        depth, conf, heat = _guided(model, xg, exg, instruction)
        margin = _assert_parity(depth[rows], conf[rows], heat[rows], ref, f"B=32 {instruction}")
        print(f"B=32 518^2 {instruction}: argmax {heat[rows].argmax(-1).tolist()} oracle top-1/top-2 margin {margin:.2e}")


def test_high_resolution_parity(model, sd):
    """Config 5: 1036 x 1036 (g = 74, 5477 tokens): bicubic position-embedding interpolation, 43 query tiles with a
    ragged tail, 86 key steps; same tolerances as at 518 (depth abs-rel 1e-2, heat-map 1e-2, argmax exact) for ALL 9
    instructions — `bottom-left` and `top-right` are the exact-tie cells of the mask at g = 74, where
    only the base attention separates the two best cells; the oracle's margin is printed."""
    x = orc.synthetic_images(1, 1036)
    ex = orc.synthetic_exif(1)
    tokens = orc.dinov2_tokens(sd, x)
    tok = model.backbone_tokens(x.cuda()).cpu()
    assert ((tok - tokens).norm() / tokens.norm()).item() < 1.5e-2
    for instruction in orc.INSTRUCTIONS:
        torch.manual_seed(11)
        ref = orc.forward_with_guidance(sd, None, ex, instruction, tokens=tokens, update_history=False)
        depth, conf, heat = _guided(model, x.cuda(), _exif(ex), instruction)
        margin = _assert_parity(depth, conf, heat, ref, f"1036^2 {instruction}")
        print(f"1036^2 {instruction}: argmax {heat.argmax(-1).tolist()} oracle top-1/top-2 margin {margin:.2e}")


def test_batch64_image_matches_oracle(model, sd):
    """Config 4's per-GPU shape (64 images of 518 x 518 in one call): image 40 of the batch against the CPU oracle."""
    B, rows = 64, [40]
    x = orc.synthetic_images(B, 518, seed=77)
    ex = orc.synthetic_exif(B, seed=78)
    tokens = orc.dinov2_tokens(sd, x[rows])
    torch.manual_seed(11)
    ref = orc.forward_with_guidance(sd, None, {k: v[rows] for k, v in ex.items()}, "top-left", tokens=tokens,
                                    update_history=False, rng_rows=(B, rows))
    depth, conf, heat = _guided(model, x.cuda(), _exif(ex), "top-left")
    _assert_parity(depth[rows], conf[rows], heat[rows], ref, "B=64 top-left")


def test_batch64_runs(model):
    """Config 4 (64 images / GPU at 518 x 518): workspace sizing and tile counts at M = 87 680 rows."""
    B = 64
    x = orc.synthetic_images(B, 518, seed=77).cuda()
    ex = _exif(orc.synthetic_exif(B, seed=78))
    depth, conf, heat = _guided(model, x, ex, "top-left")
    assert torch.isfinite(depth).all() and torch.allclose(heat.sum(-1), torch.ones(B, device=heat.device), atol=1e-5)
    d2, _, h2 = _guided(model, x[40:42], {k: v[40:42] for k, v in ex.items()}, "top-left", replay_batch=B)
    assert torch.equal(d2, depth[40:42]) and torch.equal(h2, heat[40:42])


def test_per_image_instructions_in_one_batch(model):
    """one batch, one instruction PER IMAGE (demo.py:406-432 loops over single-image calls) ==
    the rows of the nine single-instruction batched calls, bit for bit."""
    B = 9
    x = orc.synthetic_images(B, 224).cuda()
    ex = _exif(orc.synthetic_exif(B))
    instr = list(orc.INSTRUCTIONS)
    depth, conf, heat = _guided(model, x, ex, instr)
    for i, one in enumerate(instr):
        d1, c1, h1 = _guided(model, x, ex, one)
        assert torch.equal(d1[i], depth[i]) and torch.equal(c1[i], conf[i]) and torch.equal(h1[i], heat[i]), one
    assert len({int(a) for a in heat.argmax(-1)}) > 1
    with pytest.raises(ValueError):
        model.forward_with_guidance(x, ex, instr[:4])


def test_focus_map_of_last_forward(model):
    """model.focus_map(): the overlay demo.py renders from get_attention_weights(), computed on the GPU."""
    x = orc.synthetic_images(2, 224).cuda()
    ex = _exif(orc.synthetic_exif(2))
    _, _, heat = _guided(model, x, ex, "right")
    fm = model.focus_map((200, 320))
    want = orc.focus_map(heat.cpu(), 200, 320)
    assert fm.shape == (2, 200, 320)
    assert (fm.cpu() - want).abs().max().item() < 2e-6
    with pytest.raises(ValueError):
        model.focus_map((10, 10), attention=torch.zeros(1, 10).cuda())


def test_demo_single_image_path(model, sd):
    """BASELINE.json configs[0]: demo.py's single-image call — a 480 x 640 uint8 image, Resize((224, 224)) as
    experiment_B.yaml:87 / demo.py:154 configure it, instruction 'center', batch of ONE — GPU preprocessing + forward
    against the oracle's preprocessing + forward."""
    import numpy as np
    img = np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8)  # stands in for the absent 1.jpg
    x_ref = orc.demo_preprocess(img, 224).unsqueeze(0)
    x_gpu = model.preprocess(torch.from_numpy(img).unsqueeze(0), 224)
    assert torch.equal(x_gpu.cpu(), x_ref)
    ex = {"focal_length": torch.tensor([50.0]), "aperture": torch.tensor([2.8]), "iso": torch.tensor([100.0]),
          "camera_idx": torch.tensor([0])}  # demo.py:263-277 defaults
    torch.manual_seed(11)
    ref = orc.forward_with_guidance(sd, x_ref, ex, "center", update_history=False)
    depth, conf, heat = _guided(model, x_gpu, _exif(ex), "center")
    assert depth.shape == (1, 1) and heat.shape == (1, 256)
    assert ((depth.cpu() - ref["depth"]).abs() / ref["depth"].abs()).max().item() <= 1e-2
    assert (conf.cpu() - ref["confidence"]).abs().max().item() <= 1e-2
    assert (heat.cpu() - ref["heatmap"]).abs().max().item() <= 1e-2
    assert torch.equal(heat.cpu().argmax(-1), ref["heatmap"].argmax(-1))
    # demo.py:355-356 reads the scalars back
    assert abs(depth.squeeze().cpu().item() - ref["depth"].item()) / ref["depth"].item() <= 1e-2


@pytest.mark.parametrize("S,B", [(56, 3), (70, 5), (126, 1)])
def test_tiny_grids_and_odd_batches(model, sd, S, B):
    """Smallest supported grids (4 x 4, 5 x 5, 9 x 9 patches: 17 / 26 / 82 tokens — every tile is ragged) and odd batch
    sizes against the oracle."""
    x = orc.synthetic_images(B, S, seed=S)
    ex = orc.synthetic_exif(B, seed=S + 1)
    torch.manual_seed(11)
    ref = orc.forward_with_guidance(sd, x, ex, "top-right", update_history=False)
    depth, conf, heat = _guided(model, x.cuda(), _exif(ex), "top-right")
    assert ((depth.cpu() - ref["depth"]).abs() / ref["depth"].abs()).max().item() <= 1e-2
    assert (heat.cpu() - ref["heatmap"]).abs().max().item() <= 1e-2
    assert torch.equal(heat.cpu().argmax(-1), ref["heatmap"].argmax(-1))
    ref_u = orc.forward_unguided(sd, x, ex, update_history=False)
    d2, _, a2 = model(x.cuda(), _exif(ex), return_attention=True)
    assert ((d2.cpu() - ref_u["depth"]).abs() / ref_u["depth"].abs()).max().item() <= 1e-2
    assert (a2.cpu() - ref_u["heatmap"]).abs().max().item() <= 1e-2
