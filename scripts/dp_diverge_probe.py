"""Data-parallel consistency probe: N ranks (torchrun), different clips and masks per rank, the bucketed all-reduce
attached; after every training step the bit patterns of every parameter tensor are compared across ranks and the tensors
that differ are named (PROBE_SYNC=0: only after the last step, so that nothing synchronises the ranks in between).
    torchrun --nproc-per-node N scripts/dp_diverge_probe.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import wavjepa_b200 as w  # noqa: E402
from bench import CLIP_LEN, CLIPS, CROPS, MASKER, build_model  # noqa: E402
from wavjepa_b200.dist import BucketedAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
model = build_model(dev)
model.global_step = 1000
model.attach_data_parallel(BucketedAllReduce())
T = model.total_patches
masker = w.TimeInverseBlockMasker(**MASKER, channel_based_masking=False, seed=1234, row0=rank * (1 << 24), device=dev)
B = CLIPS * CROPS
clips = torch.randn(CLIPS, 1, CLIP_LEN, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
names = list(model._train_names)
report = []
for s in range(steps):
    ctx, tgt, vis = masker(batch_size=B, n_times=T, in_channels=1)
    x16 = model.on_after_batch_transfer((clips, ctx.view(CLIPS, CROPS, T), tgt.view(CLIPS, CROPS, -1, T), vis.view(CLIPS, CROPS, -1, T)), 0)[0]
    loss = model.train_step(x16, ctx, tgt, vis)
    if os.environ.get("PROBE_SYNC", "1") == "0" and s + 1 < steps:
        continue     # compare only after the last step: no collective / host sync between the steps, like bench.py
    # per-tensor checksums of parameters AND of the (all-reduced) gradient buffer
    cp = torch.stack([model._view(model._flat_p, n).view(torch.int32).long().sum() for n in names])
    cg = torch.stack([model._view(model._flat_g, n).view(torch.int32).long().sum() for n in names])
    allp = [torch.empty_like(cp) for _ in range(world)]
    allg = [torch.empty_like(cg) for _ in range(world)]
    dist.all_gather(allp, cp)
    dist.all_gather(allg, cg)
    if rank == 0:
        P, G = torch.stack(allp), torch.stack(allg)
        bad_p = [names[i] for i in torch.nonzero((P != P[0]).any(0)).flatten().tolist()]
        bad_g = [names[i] for i in torch.nonzero((G != G[0]).any(0)).flatten().tolist()]
        report.append({"step": s, "loss_rank0": float(loss), "params_differ": len(bad_p), "grads_differ": len(bad_g),
                       "first_params": bad_p[:8], "first_grads": bad_g[:8]})
if rank == 0:
    print(json.dumps({"world": world, "pair_grads_env": os.environ.get("WJ_GEMM_PAIR_GRADS"), "steps": report}))
dist.barrier()
dist.destroy_process_group()
