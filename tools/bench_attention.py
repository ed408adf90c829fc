import torch, sys
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops
B,T,H=32,1370,12
qkv=(torch.randn(B*T,3*H*64,device='cuda')).bfloat16()
out=torch.empty(B*T,H*64,device='cuda',dtype=torch.bfloat16)
for _ in range(3): ops.attention(qkv,out,B,T,H)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attention(qkv,out,B,T,H)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/10
print('attention ms', ms, 'TF/s', 4*B*H*T*T*64/ms/1e9)
