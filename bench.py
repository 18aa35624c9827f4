#!/usr/bin/env python
"""WavJEPA-base pre-training throughput on B200 (BASELINE.json: "train 2s-instances/s at 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (N>1: launched under torchrun)
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]   # the reference's CPU path (oracle port)

One step = one full SSL pre-training step of configs[1]: 64 clips x 8 random crops of 2.01 s (512 instances / GPU) of
synthetic 16 kHz noise, random-init WavJEPA-base weights: GPU mask generation (AudioSet masker) -> crop + normalise ->
conv encoder -> student on visible tokens -> predictor -> EMA-teacher targets -> masked latent MSE -> hand-written
backward (bucketed NCCL all-reduce overlapped when N > 1) -> EMA -> global-norm clip + AdamW.  Nothing is skipped.

`value`  : instances/s with the step's clips already resident in HBM (CUDA events, max over ranks).
`e2e`    : the same step driven from pinned HOST clips (H2D copy of the [64,1,160000] fp32 batch and a D2H read of
           the loss inside the timed region, wall clock between device synchronisations, max over ranks).
`roofline`: the dominant kernel (the tcgen05 GEMM / implicit-GEMM-conv kernel, all its launches of one step):
           executed GEMM FLOPs / summed CUDA-event durations, against the measured bf16 peak.
`cpu_baseline`: the CPU oracle port of the same step on a bounded sample (rank 0, N = 1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# activations of consecutive steps differ in size (mask-dependent token counts): let the caching allocator grow its
# segments in place instead of cudaMalloc-ing new ones in the middle of a timed step
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "WavJEPA-base train 2s-instances/s at 1/2/4/8 B200; tensor-pipe % of peak"
UNIT = "instances/s"
CLIPS, CROPS, CLIP_LEN = 64, 8, 160000
MASKER = dict(target_masks_per_context=4, context_mask_prob=0.65, context_mask_length=10, target_prob=0.25,
              target_length=10, ratio_cutoff=0.1)   # configs/masker/AudioSet.yaml


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


# ===================================================================================================== FLOP model
def conv_macs(spec, length):
    macs, cin, L, first = 0, 1, length, 0
    for i, (dim, k, s) in enumerate(spec):
        L = (L - k) // s + 1
        m = L * dim * cin * k
        if i == 0:
            first = m
        macs += m
        cin = dim
    return macs, first


def stack_macs(n, d, layers=12):
    return layers * (n * 12 * d * d + 2 * n * n * d)


def algorithmic_flops(n_c, n_v, n_t, B, T=200, D=768, Dp=384):
    """useful (mask-exact) FLOPs of one training step; masked-out work earns no credit."""
    spec = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)]
    conv, conv0 = conv_macs(spec, 32159)
    conv, conv0 = conv * B, conv0 * B
    mapper = B * T * 512 * D
    student = sum(stack_macs(int(n), D) for n in n_c)
    teacher = B * stack_macs(T, D)
    predictor = sum(stack_macs(int(n), Dp) for n in n_v)
    d2e = int(sum(n_t)) * Dp * D
    e2d = int(sum(n_c)) * D * Dp
    return 2.0 * (3 * (conv + mapper + student + e2d + predictor + d2e) - conv0 + teacher)


# ===================================================================================================== clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ===================================================================================================== our arm
def build_model(device):
    import torch
    import wavjepa_b200 as w

    torch.manual_seed(0)
    ex = w.ConvFeatureExtractor(conv_layers_spec=[(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)], in_channels=1)
    model = w.JEPA(feature_extractor=ex, transformer_encoder_cfg=w.TransformerEncoderCFG.create(),
                   transformer_encoder_layers_cfg=w.TransformerLayerCFG.create(),
                   transformer_decoder_cfg=w.TransformerEncoderCFG.create(),
                   transformer_decoder_layers_cfg=w.TransformerLayerCFG.create(d_model=384),
                   lr=4e-4, adam_betas=(0.9, 0.98), adam_weight_decay=0.04, process_audio_seconds=2.01,
                   nr_samples_per_audio=CROPS, average_top_k_layers=8)   # configs/trainer/default_trainer.yaml
    return model.to(device)


def run_native(args):
    import torch
    import torch.distributed as dist
    import wavjepa_b200 as w
    from wavjepa_b200 import _lib
    from wavjepa_b200.dist import BucketedAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model = build_model(dev)
    model.reserve_workspace(80 << 30)   # setup: the activation pool of a 512-instance step (53-66 GB) exists up front
    model.global_step = 1000          # lr(0) == 0 (warm-up from 0): start inside the warm-up so AdamW moves weights
    if world > 1:
        model.attach_data_parallel(BucketedAllReduce())
    masker = w.TimeInverseBlockMasker(**MASKER, seed=1234, row0=rank * (1 << 24), device=dev)
    B = CLIPS * CROPS
    T = model.total_patches
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_pool = 3
    host_clips = [torch.randn(CLIPS, 1, CLIP_LEN, generator=torch.Generator().manual_seed(100 * rank + i)).pin_memory()
                  for i in range(n_pool)]
    dev_clips = [c.to(dev) for c in host_clips]
    stage = torch.empty(CLIPS, 1, CLIP_LEN, device=dev)
    last_mi = {}

    def step(clips):
        ctx, tgt, vis = masker(batch_size=B, n_times=T, in_channels=1)
        starts = torch.randint(0, CLIP_LEN - model.target_length + 1, (CLIPS, CROPS), device=dev, generator=gen)
        x16, _, _, _ = model.on_after_batch_transfer((clips, ctx.view(CLIPS, CROPS, T), tgt.view(CLIPS, CROPS, 4, T),
                                                      vis.view(CLIPS, CROPS, 4, T)), 0, starts=starts)
        loss = model.train_step(x16, ctx, tgt, vis)
        last_mi["mi"] = model._last_mi
        return loss

    def note(msg):
        if os.environ.get("WJ_BENCH_VERBOSE"):
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------------------------------------------------------- device-resident arm (`value`)
    note("model built")
    for i in range(args.warmup):
        step(dev_clips[i % n_pool])
        note(f"warm-up step {i} issued")
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    k0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(args.steps):
        loss = step(dev_clips[i % n_pool])
    e1.record()
    barrier()
    launches = _lib.kernel_launches() - k0
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    final_loss = loss.item()
    note(f"timed region done: {ms_step:.1f} ms/step")

    # ---------------------------------------------------------------- end-to-end arm (`e2e`): host clips in, loss out
    def e2e_step(i):
        stage.copy_(host_clips[i % n_pool], non_blocking=True)
        return step(stage).item()

    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * args.steps / e2e_s
    note("e2e region done")
    h2d = CLIPS * CLIP_LEN * 4
    d2h = 4 + 8 * 4   # loss scalar + the mask totals read by the host to size the packed buffers

    # ---------------------------------------------------------------- per-kernel attribution (outside the timed region)
    # every rank runs this extra step (its gradient all-reduce is a collective); only rank 0 brackets its launches
    kp = None
    if rank == 0:
        with _lib.KernelProfile() as kp:
            step(dev_clips[0])
    else:
        step(dev_clips[0])
    barrier()

    out = None
    if rank == 0:
        peaks = load_peaks()
        summ = kp.summary()
        gemm_names = ("wj_gemm_bf16", "wj_gemm_dgrad_bf16", "wj_gemm_wgrad_bf16")
        g_calls = sum(summ[n][0] for n in gemm_names if n in summ)
        g_ms = sum(summ[n][1] for n in gemm_names if n in summ)
        g_flops = sum(m[0] for (n, _, _, m) in kp.records if n in gemm_names and m)
        fam_ms = {n[3:]: round(t, 3) for n, (c, t) in sorted(summ.items(), key=lambda kv: -kv[1][1])}
        fam_calls = {n[3:]: c for n, (c, t) in summ.items()}
        prof_total = sum(t for (_, t) in summ.values())
        achieved = g_flops / (g_ms * 1e-3) / 1e12
        mi = last_mi["mi"]
        f_alg = algorithmic_flops(mi.n_c.tolist(), mi.n_v.tolist(), mi.n_t.tolist(), B)
        roofline = {"bound": "tensor", "kernel": "gemm_kernel<BN,MODE> (tcgen05.mma + TMA; Linear fwd/dgrad/wgrad and "
                    "implicit-GEMM Conv1d)", "achieved": round(achieved, 1), "peak": peaks["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": round(achieved / peaks["bf16_sustained"], 4),
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                    "launches_per_step": g_calls, "avg_launch_ms": round(g_ms / max(g_calls, 1), 4),
                    "flops_per_step": g_flops, "share_of_step": round(g_ms / prof_total, 4), "traffic": None}
        tr = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
        if os.path.exists(tr):   # dram__bytes_read.sum + dram__bytes_write.sum over the GEMM launches of one step (ncu)
            t = json.load(open(tr))
            roofline["traffic"] = t["dram_bytes_per_launch_avg"]
            roofline["traffic_note"] = ("average DRAM bytes per GEMM launch, ncu over the %d launches of one step "
                                        "(profiles/r01_gemm_traffic.json); algorithmic bytes/launch (A + W + outputs) "
                                        "are of the same order: the kernel is tensor-bound, not HBM-bound" % t["launches"])
        out = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": "configs[1]: WavJEPA-base SSL pre-training step, 64 clips x 8 crops of 2.01 s "
                                      "(512 instances/GPU), random-init weights, AudioSet masker, AdamW+EMA included",
                          "instances_per_gpu": B, "global_instances": world * B, "tokens_per_instance": T,
                          "parallelism": f"dp{world}",
                          "l2": "no explicit flush: every step streams > 30 GB of activations (>> 126 MB L2) and "
                                "rotates 3 different 41 MB clip batches"},
               "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_s / args.steps * 1e3, 3)},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
               "step_tensor": {"alg_flops_per_step": f_alg, "alg_tflops": round(f_alg / (ms_step * 1e-3) / 1e12, 1),
                               "frac_of_peak": round(f_alg / (ms_step * 1e-3) / 1e12 / peaks["bf16_sustained"], 4),
                               "note": "useful mask-exact FLOPs (SURVEY.md 8d) / whole step time incl. optimizer"},
               "kernel_ms_per_step": fam_ms, "kernel_calls_per_step": fam_calls, "loss": final_loss,
               "host_cpus": os.cpu_count()}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(budget_s=25.0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out), flush=True)


# ===================================================================================================== CPU arms
def _cpu_train_setup(n_inst):
    import torch
    from oracle import inputs as oi
    from oracle import jepa_oracle as jo

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = jo.Cfg()
    sd = jo.make_state_dict(cfg, seed=0)
    names = [k for k in sd if not k.startswith(("teacher_encoder.", "pos_encoding"))]
    params = [sd[k].requires_grad_(True) for k in names]
    opt = torch.optim.AdamW(params, lr=4e-4 * 1000 / 100000, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.04)
    inp = oi.training_inputs(cfg, 1, n_inst, seed=1234, masker="audioset")

    def step():
        opt.zero_grad(set_to_none=True)
        out = jo.forward(inp["audio"], inp["ctx_masks"], inp["target_indices"], inp["ctx_and_target_masks"], sd, cfg)
        out["loss"].backward()
        with torch.no_grad():
            jo.ema_update(sd, step=1000)
        torch.nn.utils.clip_grad_norm_(params, 5.0)
        opt.step()
        return out["loss"].item()

    return step


def cpu_baseline(budget_s: float):
    """The CPU oracle port of the same training step (fp32, dense like the reference) on a bounded sample."""
    n_inst = 8
    step = _cpu_train_setup(n_inst)
    t0 = time.perf_counter()
    step()                                    # warm-up (also sizes the sample)
    t1 = time.perf_counter() - t0
    reps = max(1, min(3, int(budget_s / max(t1, 1e-3)) - 1))
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    import torch
    return {"value": round(n_inst / dt, 3), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} x one oracle training step (fwd+bwd+EMA+clip+AdamW, fp32, torch CPU) on 1 clip x "
                      f"{n_inst} crops of 2.01 s; {dt:.2f} s/step"}


def run_reference(args):
    """`--impl reference`: the reference's CPU path for the same step.  The reference is pure Python on PyTorch and
    cannot travel to the GPU box, so this arm times its op-for-op restatement (oracle/jepa_oracle.py, pinned against
    the executed reference by tests/test_oracle_cpu.py) on all host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    n_inst = 8
    step = _cpu_train_setup(n_inst)
    t0 = time.perf_counter()
    step()
    t_first = time.perf_counter() - t0
    if t_first * (args.steps + args.warmup) > 240 and n_inst > 2:   # keep the whole run within a few minutes
        n_inst = max(2, int(n_inst * 240 / (t_first * (args.steps + args.warmup))))
        step = _cpu_train_setup(n_inst)
    for _ in range(max(0, args.warmup - 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n_inst * args.steps / dt
    sample = (f"each step = one oracle training step (fwd+bwd+EMA+clip+AdamW, fp32 torch CPU, dense tokens like the "
              f"reference) on 1 clip x {n_inst} crops of 2.01 s")
    out = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[1] step on a bounded CPU sample: " + sample, "instances_per_step": n_inst},
           "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": sample},
           "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "host_cpus": os.cpu_count()}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
