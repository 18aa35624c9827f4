"""Per-launch attribution of one training step (CUDA events around every C-ABI call), grouped by entry point and GEMM
shape.  Run on the GPU box:  python scripts/profile_step.py [--clips 64] > gpurun_out/step_profile.txt"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import wavjepa_b200 as w  # noqa: E402
from wavjepa_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--crops", type=int, default=8)
ap.add_argument("--cprofile", action="store_true")
ap.add_argument("--ncu", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
model.global_step = 1000
B, T = args.clips * args.crops, 200
masker = w.TimeInverseBlockMasker(**bench.MASKER, seed=1234, row0=0, device=dev)
clips = torch.randn(args.clips, 1, bench.CLIP_LEN, device=dev)


def step():
    ctx, tgt, vis = masker(batch_size=B, n_times=T, in_channels=1)
    starts = torch.randint(0, bench.CLIP_LEN - model.target_length + 1, (args.clips, args.crops), device=dev)
    x16, _, _, _ = model.on_after_batch_transfer((clips, ctx.view(args.clips, args.crops, T),
                                                  tgt.view(args.clips, args.crops, 4, T),
                                                  vis.view(args.clips, args.crops, 4, T)), 0, starts=starts)
    return model.train_step(x16, ctx, tgt, vis)


if "--ncu" in sys.argv:   # two plain steps for an ncu capture (skip the first step's launches)
    step()
    step()
    torch.cuda.synchronize()
    sys.exit(0)
import time  # noqa: E402
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    print(f"warm-up step {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print(f"step (unprofiled): {e0.elapsed_time(e1) / 3:.2f} ms")
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host issue time of one step: {(t1 - t0) * 1e3:.1f} ms; until GPU idle: {(t2 - t0) * 1e3:.1f} ms")
if "--cprofile" in sys.argv:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
with _lib.KernelProfile() as kp:
    step()
torch.cuda.synchronize()
groups = {}
for name, a, b, meta, nbytes in kp.records:
    key = (name, meta[1:] if meta else None)
    c, t, f, nb = groups.get(key, (0, 0.0, 0.0, 0))
    groups[key] = (c + 1, t + a.elapsed_time(b), f + (meta[0] if meta else 0.0), nb + (nbytes or 0))
tot = sum(t for (_, t, _, _) in groups.values())
print(f"sum of bracketed launches: {tot:.2f} ms")
print(f"{'entry':28s} {'(M, N, K, act, f32out)':34s} {'calls':>5s} {'ms':>9s} {'%':>6s} {'TFLOP/s':>8s} {'GB/s':>7s}")
for (name, shape), (c, t, f, nb) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    tf = f / (t * 1e-3) / 1e12 if f else 0.0
    gbs = nb / (t * 1e-3) / 1e9 if nb else 0.0
    print(f"{name[3:]:28s} {str(shape) if shape else '':34s} {c:5d} {t:9.3f} {100 * t / tot:6.2f} {tf:8.1f} {gbs:7.0f}")
