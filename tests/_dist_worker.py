"""Worker of tests/test_gpu_dist.py (one process per GPU, launched by torch.distributed.run): data-parallel correctness of
the overlapped bucketed all-reduce on real NCCL (reference: Lightning DDP, train.py:174-179).  Prints DIST_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import wavjepa_b200 as w  # noqa: E402
from wavjepa_b200.dist import BucketedAllReduce  # noqa: E402

SPEC = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)]


class CheckingAllReduce(BucketedAllReduce):
    """Keeps a copy of every bucket's LOCAL gradients before they are summed in place."""

    def begin(self, flat):
        super().begin(flat)
        self.local = []

    def _launch(self, lo):
        if lo < self._hi:
            torch.cuda.current_stream().synchronize()   # (test only: the slice is final, copy it before NCCL sums into it)
            self.local.append((lo, self._hi, self._flat[lo:self._hi].clone()))
        super()._launch(lo)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(100 + rank)          # ranks start DIFFERENT on purpose: attach_data_parallel must repair that
    ex = w.ConvFeatureExtractor(conv_layers_spec=SPEC, in_channels=1)
    model = w.JEPA(feature_extractor=ex, transformer_encoder_cfg=w.TransformerEncoderCFG.create(),
                   transformer_encoder_layers_cfg=w.TransformerLayerCFG.create(),
                   transformer_decoder_cfg=w.TransformerEncoderCFG.create(),
                   transformer_decoder_layers_cfg=w.TransformerLayerCFG.create(d_model=384), lr=4e-4,
                   adam_weight_decay=0.04, process_audio_seconds=2.01, nr_samples_per_audio=4,
                   average_top_k_layers=8).to(dev)
    model.global_step = 50000 + 7 * rank
    red = CheckingAllReduce(bucket_bytes=8 << 20)     # small buckets: many overlapped launches
    model.attach_data_parallel(red)

    def gathered(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    # (0) rank 0's state everywhere after attach
    assert model.global_step == 50000, model.global_step
    for t in (model._flat_p, model._flat_t):
        g = gathered(t)
        assert all(torch.equal(g[0], x) for x in g), "parameters differ across ranks after attach_data_parallel"

    B, T = 8, 200
    masker = w.TimeInverseBlockMasker(4, 0.65, 10, 0.25, 10, 0.1, seed=5, row0=rank * (1 << 20), device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)                         # every rank sees its own data (C3)
    for step in range(3):
        audio = torch.randn(B, 1, model.target_length, device=dev, generator=gen).bfloat16()
        ctx, tgt, vis = masker(batch_size=B, n_times=T, in_channels=1)
        loss = model.train_step(audio, ctx, tgt, vis)
        assert torch.isfinite(loss).item()
        # (1) every bucket of the all-reduced gradient == the sum of the ranks' local gradients of that bucket
        assert len(red.local) >= 10 and len(red.local) == len(red.launched)
        covered = 0
        for lo, hi, loc in red.local:
            parts = gathered(loc)
            ref = torch.stack([p.double() for p in parts]).sum(0)
            got = model._flat_g[lo:hi].double()
            err = (got - ref).norm() / (ref.norm() + 1e-30)
            assert err.item() < 1e-6, (step, lo, hi, err.item())
            covered += hi - lo
        assert covered == model._flat_g.numel()
        # (2) replicas stay bit-identical (same summed gradients, same fused optimizer + EMA)
        for t in (model._flat_p, model._flat_t, model._adam_m, model._adam_v, model._flat_w16.view(torch.uint8)):
            g = gathered(t)
            assert all(torch.equal(g[0], x) for x in g), f"replicas diverged at step {step}"
    # ranks trained on different data: the local gradients must really differ (the check above is not vacuous)
    parts = gathered(red.local[0][2])
    assert not torch.equal(parts[0], parts[-1])
    dist.barrier()
    if rank == 0:
        print("DIST_OK", world, "ranks,", len(red.local), "buckets per step")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
