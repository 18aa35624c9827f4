"""DRAM traffic per kernel family of ONE training step, from an ncu launch list of this build:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/traffic.csv python scripts/profile_step.py --ncu
    python scripts/kernel_traffic.py gpurun_out/traffic.csv > profiles/r02_kernel_traffic.json

profile_step.py --ncu runs two steps; the launches of the LAST step are used (the first step also allocates).  bench.py
reads the result for `roofline.traffic` (the GEMM family) and for the `traffic` field of its `hbm_kernels` entries."""
import collections
import csv
import json
import re
import sys

# kernel (C++ name prefix) -> the C-ABI entry point (without wj_) whose launches it serves
FAMILY = [("gemm_pair_kernel", "gemm"), ("gemm_kernel", "gemm"), ("conv0_fwd_kernel", "conv0_gn_gelu_fwd"),
          ("conv0_moments_kernel", "conv0_gn_gelu_fwd"), ("conv0_stats_kernel", "conv0_gn_gelu_fwd"),
          ("conv0_bwd_kernel", "conv0_gn_gelu_bwd"), ("conv0_bwd_finalize_kernel", "conv0_gn_gelu_bwd"),
          ("layernorm_fwd_kernel", "add_layernorm_fwd"), ("layernorm_bwd_kernel", "add_layernorm_bwd"),
          ("target_accum_kernel", "target_accum"), ("instance_stats_kernel", "target_accum"),
          ("masked_mse_kernel", "masked_mse"), ("adamw_kernel", "adamw_ema_step"), ("sumsq_kernel", "sumsq"),
          ("colsum8_kernel", "colsum"), ("colsum_kernel", "colsum"), ("crop_norm_kernel", "crop_norm"),
          ("attn_fwd_tc_kernel", "attn_varlen_fwd"), ("attn_fwd_kernel", "attn_varlen_fwd"),
          ("attn_bwd_tc_kernel", "attn_varlen_bwd_bias"), ("attn_bwd_kernel", "attn_varlen_bwd_bias")]


def family(name):
    base = re.sub(r"^void\s+", "", name).replace("wj::", "")
    for k, f in FAMILY:
        if base.startswith(k):
            return f
    return re.sub(r"[<(].*", "", base)


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    iid, iname, imet, iunit, ival = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= ival or not r[iid].isdigit():
            continue
        d = launches.setdefault(int(r[iid]), {"name": r[iname]})
        v = float(r[ival].replace(",", ""))
        u = r[iunit].lower()
        if r[imet] == "gpu__time_duration.sum":
            d["us"] = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
        else:
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
            d[r[imet]] = v * scale
    ids = sorted(launches)
    # the last step starts at the last masks_kernel launch (first kernel of a step)
    starts = [i for i in ids if "masks_kernel" in launches[i]["name"]]
    first = starts[-1] if starts else ids[0]
    fam = collections.OrderedDict()
    for i in ids:
        if i < first:
            continue
        d = launches[i]
        f = fam.setdefault(family(d["name"]), {"launches": 0, "time_us_serialized": 0.0, "dram_bytes_per_step": 0.0})
        f["launches"] += 1
        f["time_us_serialized"] += d.get("us", 0.0)
        f["dram_bytes_per_step"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    for f in fam.values():
        f["dram_bytes_per_launch_avg"] = f["dram_bytes_per_step"] / max(f["launches"], 1)
        f["time_us_serialized"] = round(f["time_us_serialized"], 1)
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                     "the launches of one training step (python scripts/profile_step.py --ncu), scripts/kernel_traffic.py",
           "launches_in_step": sum(f["launches"] for f in fam.values()), "gemm": fam.get("gemm"), "by_entry": fam}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
