import torch
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
n=1<<30
a=torch.empty(n,dtype=torch.uint8,device='cuda'); b=torch.empty(n,dtype=torch.uint8,device='cuda')
af=a.view(torch.float32); bf=b.view(torch.float32)
ms=t(lambda: a.zero_()); print(f"memset 1 GiB: {ms:.3f} ms {n/ms/1e6:.0f} GB/s write")
ms=t(lambda: bf.copy_(af)); print(f"copy 1 GiB: {ms:.3f} ms {2*n/ms/1e6:.0f} GB/s r+w")
ms=t(lambda: af.sum()); print(f"sum 1 GiB: {ms:.3f} ms {n/ms/1e6:.0f} GB/s read")
ms=t(lambda: torch.add(af,bf,out=bf)); print(f"add 2r1w: {ms:.3f} ms {3*n/ms/1e6:.0f} GB/s")
h=a.view(torch.bfloat16)
ms=t(lambda: h.fill_(1.0)); print(f"fill bf16 1 GiB: {ms:.3f} ms {n/ms/1e6:.0f} GB/s write")
