import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, torch
import torch.nn.functional as F
from wavjepa_b200 import ops
DEV = "cuda"
torch.manual_seed(0)

def poison():
    big = [torch.full((256 << 20,), float("nan"), device=DEV) for _ in range(8)]   # 8 GiB of NaN, then free
    del big

B, Cin, L, C = 4, 1, 32159, 512
x = torch.randn(B, Cin, L, device=DEV).bfloat16()
w = torch.randn(C, Cin, 10, device=DEV) * math.sqrt(2.0 / 10)
gamma = 1.0 + 0.1 * torch.randn(C, device=DEV); beta = 0.1 * torch.randn(C, device=DEV)
L_out = (L - 10) // 5 + 1
dy = (torch.randn(B, L_out, C, device=DEV) * 1e-6).bfloat16()
outs = []
for trial in range(3):
    poison()
    out = torch.empty(B, L_out, C, device=DEV, dtype=torch.bfloat16)
    mom, stats, red = ops.conv0_workspaces(B, Cin, C, DEV, backward=True)
    dge = torch.empty_like(out)
    ops.conv0_fwd(x, w, gamma, beta, out, mom, stats, dge)
    dw = torch.zeros_like(w); dg = torch.zeros(C, device=DEV); db = torch.zeros(C, device=DEV)
    ops.conv0_bwd(x, w, gamma, beta, mom, stats, dy, dge, red, dw, dg, db)
    torch.cuda.synchronize()
    outs.append((out.float().clone(), dw.clone(), dg.clone(), db.clone()))
for i in (1, 2):
    print("trial", i, "fwd equal", torch.equal(outs[0][0], outs[i][0]),
          "dw rel", ((outs[0][1] - outs[i][1]).norm() / outs[0][1].norm()).item(),
          "dg rel", ((outs[0][2] - outs[i][2]).norm() / outs[0][2].norm()).item(),
          "db rel", ((outs[0][3] - outs[i][3]).norm() / outs[0][3].norm()).item(), "nan", torch.isnan(outs[i][1]).any().item())
# fp64 reference
wr = w.bfloat16().double().requires_grad_(True); gr = gamma.double().requires_grad_(True); br = beta.double().requires_grad_(True)
h = F.conv1d(x.double(), wr, stride=5)
hq = h + (h.float().bfloat16().double() - h).detach()
y = F.gelu(F.group_norm(hq, C, gr, br, 1e-5))
y.backward(dy.double().transpose(1, 2))
print("vs fp64: dw rel", ((outs[0][1].double() - wr.grad).norm() / wr.grad.norm()).item(),
      "dg rel", ((outs[0][2].double() - gr.grad).norm() / gr.grad.norm()).item(),
      "db rel", ((outs[0][3].double() - br.grad).norm() / br.grad.norm()).item(),
      "| P-level cancellation: |dw| rms", outs[0][1].pow(2).mean().sqrt().item())
