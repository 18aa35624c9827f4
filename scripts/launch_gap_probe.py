import sys, os
sys.path.insert(0, '/root/repo')
import torch
from wavjepa_b200 import ops
dev='cuda'
def run(M,N,K,reps=300):
    a=torch.randn(M,K,device=dev).bfloat16(); w=(torch.randn(N,K,device=dev)*0.05).bfloat16()
    outs=[torch.empty(M,N,device=dev,dtype=torch.bfloat16) for _ in range(4)]
    for i in range(10): ops.gemm(ops.plain_operand(a), w, M, 1, outs[i%4])
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): ops.gemm(ops.plain_operand(a), w, M, 1, outs[i%4])
    e1.record(); torch.cuda.synchronize()
    print(f"M{M} N{N} K{K}: {e0.elapsed_time(e1)/reps*1e3:.2f} us per back-to-back launch")
if '--ncu' in sys.argv:
    for (M,N,K) in ((19906,768,768),(19906,2304,768),(172953,384,384)):
        a=torch.randn(M,K,device=dev).bfloat16(); w=(torch.randn(N,K,device=dev)*0.05).bfloat16(); o=torch.empty(M,N,device=dev,dtype=torch.bfloat16)
        for _ in range(3): ops.gemm(ops.plain_operand(a), w, M, 1, o)
    torch.cuda.synchronize()
else:
    run(19906,768,768); run(19906,2304,768); run(172953,384,384)
