#!/usr/bin/env python
"""Turns raw ncu output brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches.csv profiles/r1_step_launches.md
    python tools/summarize_profiles.py full     gpurun_out/x.ncu-rep    profiles/r1_x_ncu.md

`launches`: CSV of `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py; prints the last
full step (cold-cache, serialised per-launch times: compare SHARES, not absolutes).
`full`: `.ncu-rep` of `ncu --set full`; prints duration, tensor-pipe utilisation, DRAM bytes, registers.
`traffic`: `.ncu-rep` holding the distance GEMM -> profiles/roofline_traffic.json (DRAM bytes per launch + a hash of the
kernel's sources; bench.py reports `roofline.traffic` from it only while the sources are unchanged).

    python tools/summarize_profiles.py traffic  gpurun_out/assign_r2.ncu-rep profiles/roofline_traffic.json
"""
import csv
import subprocess
import sys


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    names = [(d[ki], float(d[vi].replace(",", ""))) for d in data]
    # one step = the launches between two successive frames_prepare kernels; take the last COMPLETE one
    starts = [i for i, (n, _) in enumerate(names) if "frames_prepare" in n]
    s, e = starts[-2], starts[-1]
    out, tot = [], 0.0
    for n, v in names[s:e]:
        if "FillFunctor" in n or "direct_copy" in n or "bfloat16_copy" in n or "packed_init" in n:
            continue
        out.append((n.split("(")[0].replace("void ", "").replace("pero::", ""), v / 1000.0))
        tot += v / 1000.0
    with open(dst, "w") as f:
        f.write(f"# Launch list of one bench step (ncu gpu__time_duration.sum, --clock-control none)\n\n")
        f.write(f"Source: `{src}`; command: `python bench.py --steps 2 --warmup 1 --no-graph --skip-cpu --skip-e2e --skip-configs --skip-gpu-baseline`.\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: read the SHARE column.\n\n")
        f.write("| # | kernel | us | share |\n|---|---|---|---|\n")
        for k, (n, v) in enumerate(out):
            f.write(f"| {k} | `{n[:90]}` | {v:.1f} | {100 * v / tot:.1f}% |\n")
        f.write(f"\nTotal kernel time of the step under ncu: {tot:.1f} us over {len(out)} launches.\n")
    print(f"wrote {dst}: {len(out)} launches, {tot:.1f} us")


WANT = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    ki = H.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in data:
            f.write(f"## `{r[ki][:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in H:
                    i = H.index(w)
                    f.write(f"| {w} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(f"wrote {dst}: {len(data)} kernels")


TRAFFIC_SOURCES = ["pero_pretraining_b200/csrc/gemm_core.cuh", "pero_pretraining_b200/csrc/gemm_host.cuh",
                   "pero_pretraining_b200/csrc/epilogues.cuh", "pero_pretraining_b200/csrc/ptx.cuh",
                   "pero_pretraining_b200/csrc/vq_assign.cu"]


def _bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value.replace(",", "")) * scale


def traffic(src, dst):
    import hashlib
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units, data = rows[0], rows[1], rows[2:]
    ki, ri, wi = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
    sel = [r for r in data if "ArgminEpi" in r[ki]]
    rd = sum(_bytes(r[ri], units[ri]) for r in sel) / len(sel)
    wr = sum(_bytes(r[wi], units[wi]) for r in sel) / len(sel)
    hsh = hashlib.sha256()
    for f in TRAFFIC_SOURCES:
        hsh.update(open(os.path.join(root, f), "rb").read())
    rec = {"kernel": sel[0][ki][:80], "launches": len(sel), "dram_read_bytes": rd, "dram_write_bytes": wr,
           "traffic_bytes": rd + wr, "capture": os.path.basename(src), "sources": TRAFFIC_SOURCES, "sources_sha256": hsh.hexdigest()}
    json.dump(rec, open(dst, "w"), indent=1)
    print(f"wrote {dst}: {rd / 1e6:.2f} MB read + {wr / 1e6:.2f} MB written per launch over {len(sel)} launches")


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
