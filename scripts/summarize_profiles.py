"""Turns the raw ncu outputs brought back in gpurun_out/ (scripts/gpu_profile_round.sh) into the tracked summaries
under profiles/.

  python scripts/summarize_profiles.py gpurun_out/launches_r01.csv gpurun_out/prof_conv_r01b.ncu-rep r01 <images in the run> \
         [gpurun_out/prof_wgrad_r01.ncu-rep]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_summary(path, tag, n_img):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    t = collections.defaultdict(float); n = collections.Counter(); rd = collections.defaultdict(float); wr = collections.defaultdict(float)
    ui = hdr.index("Metric Unit")
    def to_bytes(v, unit):
        u = unit.lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            t[name] += v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1); n[name] += 1
        elif r[mi] == "dram__bytes_read.sum":
            rd[name] += to_bytes(v, r[ui])
        elif r[mi] == "dram__bytes_write.sum":
            wr[name] += to_bytes(v, r[ui])
    tot = sum(t.values())
    lines = ["# ncu launch list summary (%s)" % tag, "",
             "Command: `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 1 --batch 16 "
             "--no-cpu-baseline --train-steps 1 --no-graph` (cold-cache, serialised launches: compare SHARES, not absolute times).",
             "The run executes %d images through the eval hot path (3 warm-up + 1 timed + 2 e2e + 1 roofline step of 16) and then "
             "4 joint training steps (3 warm-up + 1 timed; batch 8, 224^2 crops): `conv_wgrad_kernel`, `adam_kernel`, `prelu_*` and "
             "the `at::native` glue kernels belong to the training leg; `us / image` divides by the eval images only. DRAM columns "
             "are zero in this time-only pass (the per-kernel DRAM bytes of the conv kernel are in `%s_conv_traffic.json`, "
             "captured earlier in the round with the dram metrics enabled)." % (n_img, tag), "",
             "| kernel | launches | total ms | us / image | share | DRAM read MB / image | DRAM write MB / image |", "|---|---|---|---|---|---|---|"]
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        if v / tot < 0.001:
            continue
        lines.append("| `%s` | %d | %.2f | %.1f | %.1f %% | %.1f | %.1f |" % (k[:80], n[k], v / 1e6, v / 1e3 / n_img, 100 * v / tot,
                                                                             rd[k] / 1e6 / n_img, wr[k] / 1e6 / n_img))
    lines.append("")
    lines.append("Total kernel time: %.2f ms (%.3f ms / image)." % (tot / 1e6, tot / 1e6 / n_img))
    conv = "csbsr::conv_igemm_kernel"
    ck = [k for k in t if "conv_igemm_kernel" in k]          # both template instances (cta_group::1 / ::2)
    c_t, c_n, c_b = sum(t[k] for k in ck), sum(n[k] for k in ck), sum(rd[k] + wr[k] for k in ck)
    if c_b > 0:
        out = {"kernel": conv, "share_of_gpu_time": c_t / tot, "launches_per_image": c_n / n_img,
               "dram_bytes_per_image": c_b / n_img, "us_per_image": c_t / 1e3 / n_img}
        with open(os.path.join(ROOT, "profiles", "%s_conv_traffic.json" % tag), "w") as f:
            json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "profiles", "%s_launches.md" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


def full_summary(rep, tag, kernel="conv_igemm", cmd=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["ID", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    lines = ["# ncu --set full capture of `csbsr::%s_kernel` (%s)" % (kernel, tag), "",
             cmd or ("Command: `ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 700 -c 8 python bench.py "
                     "--steps 1 --warmup 1 --batch 8 --no-cpu-baseline --no-train` (8 consecutive conv launches of one KBPN stage)."), "",
             "| " + " | ".join("%s [%s]" % (w.replace("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %")
                                              .replace(".avg.pct_of_peak_sustained_elapsed", " %").replace("launch__", ""), units[i]) for w, i in idx) + " |",
             "|" + "---|" * len(idx)]
    for r in data:
        lines.append("| " + " | ".join(r[i] for _, i in idx) + " |")
    with open(os.path.join(ROOT, "profiles", "%s_%s_ncu_full.md" % (tag, kernel)), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[4:10]))


if __name__ == "__main__":
    launches, rep, tag, n_img = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launch_summary(launches, tag, n_img)
    if os.path.exists(rep):
        full_summary(rep, tag)
    if len(sys.argv) > 5 and os.path.exists(sys.argv[5]):
        full_summary(sys.argv[5], tag, "conv_wgrad",
                     "Command: `ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 40 -c 4 python "
                     "scripts/time_train.py --steps 1 --warmup 1` (4 consecutive weight-gradient launches of the backward pass, "
                     "batch 8, 224^2 crops).")
