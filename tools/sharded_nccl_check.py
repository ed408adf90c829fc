"""Batch-sharded inference over NCCL on N GPUs: every rank runs its contiguous shard, outputs are
all_gather'ed, and rank 0 checks them bit for bit against its own un-sharded run of the global batch.
    torchrun --nproc-per-node 2 tools/sharded_nccl_check.py"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cognitive_aim_depth_estimation_b200.model import create_model
from cognitive_aim_depth_estimation_b200.sharding import ShardedInference
from oracle import cogaim_oracle as orc

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
cfg = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
model = create_model(cfg, {"num_cameras": 71}, device="cuda")
model.load_state_dict(orc.build_state_dict(0))
B, S = 13, 224  # ragged shards
x = orc.synthetic_images(B, S).cuda()
ex = {k: v.cuda() for k, v in orc.synthetic_exif(B).items()}
runner = ShardedInference(model, rank, world, gather=True)
ok = True
for instr in ("center", "bottom-left"):
    torch.manual_seed(11)
    got = runner.forward_with_guidance(x, ex, instr, return_attention=True)
    torch.manual_seed(11)
    want = model.forward_with_guidance(x, ex, instr, return_attention=True)
    ok = ok and all(torch.equal(a, b) for a, b in zip(got, want)) and got[2].shape == (B, (S // 14) ** 2)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"sharded NCCL gather over {world} GPUs == un-sharded run: {'OK (bit-exact)' if flag.item() else 'MISMATCH'}")
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
