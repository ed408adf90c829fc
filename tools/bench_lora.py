"""Cost of the fused LoRA path (lora_mode: fused, one extra 64-wide K block on the adapted projection + the t = x A^T GEMM)
against the merged path (adapters folded into the packed weight: free) on the backbone at the benchmark shape."""
import sys
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops
from cognitive_aim_depth_estimation_b200.model import create_model
dev = torch.device('cuda:0')
B, S = 32, 518
x = [torch.randn(B, 3, S, S, device=dev) for _ in range(3)]
for mode, target in (("merge", "query"), ("fused", "query"), ("fused", "attention_output")):
    torch.manual_seed(0)
    m = create_model({"model": {}, "use_lora": True, "lora_merge_target": target, "lora_mode": mode}, {"num_cameras": 71}, device=dev)
    for i in range(4):
        m.backbone_tokens(x[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        m.backbone_tokens(x[i % 3])
    e1.record(); torch.cuda.synchronize()
    print(f"backbone B={B} S={S} lora_mode={mode:5s} target={target:16s}: {e0.elapsed_time(e1) / 30:.3f} ms / step")
    del m
    torch.cuda.empty_cache()
