"""Turns the raw ncu outputs brought back in gpurun_out/ (scripts/gpu_profile_round.sh) into the tracked summaries under profiles/.

  python scripts/summarize_profiles.py r02
reads  gpurun_out/launches_eval_<tag>.csv, launches_train_<tag>.csv   (one eval pass of 16 images / one training step)
       gpurun_out/prof_{conv,chain,hd,wgrad}_<tag>_raw.csv               (`ncu --page raw --csv` of the --set full captures)
writes profiles/<tag>_launches_eval.md, <tag>_launches_train.md, <tag>_conv_traffic.json, <tag>_{conv_igemm,chain,hd_fused,conv_wgrad}_ncu_full.md
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_bytes(v, unit):
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit.lower(), 1)


def launch_summary(path, out_md, title, per, per_name, traffic_json=None):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    t = collections.defaultdict(float); n = collections.Counter(); rd = collections.defaultdict(float); wr = collections.defaultdict(float)
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            t[name] += v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1); n[name] += 1
        elif r[mi] == "dram__bytes_read.sum":
            rd[name] += _to_bytes(v, r[ui])
        elif r[mi] == "dram__bytes_write.sum":
            wr[name] += _to_bytes(v, r[ui])
    tot = sum(t.values())
    own = sum(v for k, v in t.items() if k.startswith("csbsr::") or k.startswith("kp::") or "csbsr" in k)
    lines = ["# " + title, "",
             "`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` "
             "around exactly one pass (`scripts/ncu_one_pass.py`; cold-cache, serialised launches: compare SHARES, not absolute times).", "",
             "| kernel | launches | total ms | us / %s | share | DRAM read MB / %s | DRAM write MB / %s |" % (per_name, per_name, per_name),
             "|---|---|---|---|---|---|---|"]
    for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
        if v / tot < 0.001:
            continue
        lines.append("| `%s` | %d | %.3f | %.1f | %.1f %% | %.1f | %.1f |" % (k[:80], n[k], v / 1e6, v / 1e3 / per, 100 * v / tot,
                                                                              rd[k] / 1e6 / per, wr[k] / 1e6 / per))
    lines += ["", "Total kernel time %.2f ms over %d launches; own kernels (csbsr:: / kp::) %.1f %% of it, the rest are aten fills / copies / "
              "elementwise kernels." % (tot / 1e6, sum(n.values()), 100 * own / tot)]
    if traffic_json:
        ck = [k for k in t if "conv_igemm_kernel" in k]
        c_t, c_n, c_b = sum(t[k] for k in ck), sum(n[k] for k in ck), sum(rd[k] + wr[k] for k in ck)
        with open(traffic_json, "w") as f:
            json.dump({"kernel": "csbsr::conv_igemm_kernel", "share_of_gpu_time": c_t / tot, "launches_per_image": c_n / per,
                       "dram_bytes_per_image": c_b / per, "us_per_image": c_t / 1e3 / per,
                       "source": os.path.basename(path) + " (one eval pass of %d images, eval kernels only)" % per}, f, indent=1)
    with open(out_md, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[4:12]))


WANT = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


NOTES = {
    "chain": "Reading: DRAM traffic is at or below the algorithmic bytes of a stage (SR chain: fp32 planes in, the 16 x 448^2 x 64 bf16 = 411 MB map "
             "out, of which 357 MB reached DRAM inside the kernel's window; CAT chain: that map in, 64 floats per image out). L2->SM traffic of the "
             "CAT chain is 2.5x its DRAM reads (halo columns of neighbouring strips, 136-pixel TMA rows for 112-122 valid columns, row bands "
             "overlapping by the layers' vertical halo). The tensor pipe is 35 % active on N = 32 / 64 instructions whose floor is 40-48 cycles per "
             "MMA against 16 / 32 at full rate (`r02_ubench_mma.txt`).",
    "hd": "Reading: 24 MB of DRAM reads for 915 MB of L2->SM traffic per 16 images: every (image, threshold) item re-reads the image's per-corner "
          "threshold ranges (431 KB) and looks distances up in the gt EDT map, all L2 hits; no tensor work, warps 50 % active -- integer / latency-bound.",
    "conv": "Reading: the 8x8/s4 deconv phases (grid 148, 640 threads) show L2->SM traffic of 3-4x their DRAM bytes (weight tiles re-read per pixel "
            "tile, DESIGN.md section 4.1) with lts throughput 60-75 %; the cta_group::2 launches (conv_igemm_kernel<1>) reach the highest tensor-pipe "
            "activity.",
    "wgrad": "Reading: the first launches of the backward pass are the small segmentation-head layers; the 8x8/s4 and 224^2 layers of KBPN follow "
             "later in the step (per-layer table in `r02_train_step_kernels.txt`).",
}


def full_summary(raw_csv, out_md, title, cmd, note=None):
    rows = list(csv.reader(open(raw_csv, errors="ignore")))
    rows = [r for r in rows if r]
    hi = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    short = lambda w: (w.replace("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %")
                        .replace(".avg.pct_of_peak_sustained_elapsed", " %").replace("launch__", "")
                        .replace("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"))
    lines = ["# " + title, "", cmd, "", "| " + " | ".join("%s [%s]" % (short(w), units[i]) for w, i in idx) + " |", "|" + "---|" * len(idx)]
    for r in data:
        lines.append("| " + " | ".join((r[i].split("(")[0] if w == "Kernel Name" else r[i])[:44] for w, i in idx) + " |")
    if note:
        lines += ["", note]
    with open(out_md, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(out_md, len(data), "launches")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    go, pr = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    os.makedirs(pr, exist_ok=True)
    p = os.path.join(go, "launches_eval_%s.csv" % tag)
    if os.path.exists(p):
        launch_summary(p, os.path.join(pr, "%s_launches_eval.md" % tag), "ncu launch list of one eval pass, 16 x 448^2 images (%s)" % tag, 16, "image",
                       os.path.join(pr, "%s_conv_traffic.json" % tag))
    p = os.path.join(go, "launches_train_%s.csv" % tag)
    if os.path.exists(p):
        launch_summary(p, os.path.join(pr, "%s_launches_train.md" % tag), "ncu launch list of one joint training step, 8 x 224^2 crops (%s)" % tag, 1, "step")
    base = "`ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:%s %s python scripts/ncu_one_pass.py %s`"
    for key, name, kern, flags, what, note in (
            ("conv", "conv_igemm", "conv_igemm", "-s 20 -c 40", "eval 16", "40 consecutive conv launches of the first KBPN stages of one 16-image pass"),
            ("chain", "chain", "chain_kernel", "-c 2", "eval 16", "the fe_SR and fe_cat chains of the first KBlock stage, 16 images"),
            ("hd", "hd_fused", "hd_fused", "-c 1", "eval 16", "the fused HD / MSD sweep of 16 images x 99 thresholds"),
            ("wgrad", "conv_wgrad", "conv_wgrad_kernel", "-c 40", "train", "the first 40 weight-gradient launches of the backward pass (segmentation head first)")):
        p = os.path.join(go, "prof_%s_%s_raw.csv" % (key, tag))
        if os.path.exists(p) and os.path.getsize(p) > 100:
            full_summary(p, os.path.join(pr, "%s_%s_ncu_full.md" % (tag, name)), "ncu --set full capture of `%s` (%s)" % (kern, tag),
                         "Command: " + base % (kern, flags, what) + " (" + note + ").", NOTES.get(key))
