"""Run-to-run difference of one training step: the SAME model state, clips and masks stepped twice (two identical models),
then every parameter gradient compared.  What can differ is the order of fp32 / fp64 atomic sums (split-token weight
gradients, column sums, LayerNorm parameter gradients, conv-0 reduction); the activation gradients themselves take no
atomics.  With --det the library's deterministic mode (workspace + fixed-order reductions) is switched on: expect zeros.
    python scripts/determinism_probe.py [clips] [--det]        (default 16 clips x 8 crops = 128 instances)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402

import wavjepa_b200 as w  # noqa: E402
from bench import MASKER, build_model  # noqa: E402

dev = torch.device("cuda", 0)
det = "--det" in sys.argv
argv = [a_ for a_ in sys.argv[1:] if a_ != "--det"]
clips = int(argv[0]) if argv else 16
if det:
    from wavjepa_b200 import ops
    ops.set_deterministic(True, 256 << 20)
crops = 8
a, b = build_model(dev), build_model(dev)
b.load_state_dict(a.state_dict())
a.global_step = b.global_step = 1000
T = a.total_patches
masker = w.TimeInverseBlockMasker(**MASKER, channel_based_masking=False, seed=7, device=dev)
B = clips * crops
ctx, tgt, vis = masker(batch_size=B, n_times=T, in_channels=1)
audio = torch.randn(clips, 1, 160000, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
starts = torch.randint(0, 160000 - a.target_length, (clips, crops), device=dev,
                       generator=torch.Generator(device=dev).manual_seed(2))
batch = (audio, ctx.view(clips, crops, T), tgt.view(clips, crops, -1, T), vis.view(clips, crops, -1, T))
xa = a.on_after_batch_transfer(batch, 0, starts=starts)[0]
xb = b.on_after_batch_transfer(batch, 0, starts=starts)[0]
assert torch.equal(xa, xb)
la = a.train_step(xa, ctx, tgt, vis)
lb = b.train_step(xb, ctx, tgt, vis)
rows, worst, same = [], 0.0, 0
for n in a._train_names:
    ga, gb = a._view(a._flat_g, n).double(), b._view(b._flat_g, n).double()
    d = (ga - gb).norm().item() / max(gb.norm().item(), 1e-30)
    same += int(torch.equal(ga, gb))
    worst = max(worst, d)
    rows.append((d, n))
rows.sort(reverse=True)
print(json.dumps({"deterministic_mode": det, "instances": B, "loss_a": la.item(), "loss_b": lb.item(), "loss_abs_diff": abs(la.item() - lb.item()),
                  "parameter_tensors": len(rows), "bit_identical_gradient_tensors": same, "worst_rel_l2_diff": worst,
                  "worst_tensors": [(n, float(f"{d:.3e}")) for d, n in rows[:6]],
                  "params_after_step_max_abs_diff": max((pa.detach() - pb.detach()).abs().max().item()
                                                        for pa, pb in zip(a.parameters(), b.parameters()))}))
