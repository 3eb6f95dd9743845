#!/usr/bin/env python
"""Turn the raw GPU evidence a `gpurun` trip leaves under gpurun_out/ into the tracked summaries under
profiles/ (the .ncu-rep files themselves are too big to commit).

    python tools/summarize_profiles.py [--round r02] [--src gpurun_out/evidence] [--dst profiles]

inputs (all optional, whatever exists is summarised):
    <src>/bench.json              bench.py line of the default run          -> <round>_bench_1gpu.json
    <src>/launches.csv            ncu --metrics gpu__time_duration.sum --csv of the same bench command
                                                                            -> <round>_launches_bench_B128.csv, <round>_launch_shares.txt
    <src>/prof_{gemm16,attn16,vq,ln}.ncu-rep   ncu --set full captures of tools/ncu_target.py
                                                                            -> <round>_ncu_*.txt
"""
from __future__ import annotations

import argparse
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
]

HEADERS = {
    "gemm16": "# gemm_tc_kernel<KIND=f16, BN=256, CG=2>: to_qkv forward shape M=131072 N=2304 K=768 with fp16 output, and net.2 forward\n"
              "# shape M=131072 N=768 K=3072 with fp32 output (python tools/ncu_target.py gemm16).\n"
              "# algorithmic: 463.9 GFLOP / 0.809 GB (A 201 MB + B 3.5 MB + C 604 MB) and 618.5 GFLOP / 1.213 GB (A 805 + B 4.7 + C 403 MB).\n"
              "# traffic = dram read+write below.\n",
    "attn16": "# fp16 attention core, B=32 N=1024 heads=12 dh=64 (python tools/ncu_target.py attn16)\n"
              "# algorithmic flops: fwd 4*B*H*N*N*dh = 103 GFLOP; bwd 10*B*H*N*N*dh = 258 GFLOP (dKV 8, dQ 6 GEMM-units of 2*N*N*dh executed: 14;\n"
              "# delta = rowsum(dO*O) is computed inside the dQ kernel, which therefore runs first)\n",
    "vq": "# vq_fwd_kernel 131072 tokens x 8192 codes x 32 dims (python tools/ncu_target.py vq)\n"
          "# algorithmic: 68.7 GFLOP fp32 FMA, 34.6 MB (z in, z_q out, idx out, codebook)\n",
    "ln": "# ln_fwd / ln_bwd kernels M=131072 D=768 (python tools/ncu_target.py ln)\n"
          "# fp16 data path configuration.  algorithmic bytes: fwd 604 MB (read x fp32, write y fp16), bwd 1611 MB (read dy fp16, x, dres;\n"
          "# write dx fp32 and its scaled fp16 copy)\n",
}
STALE_NOTES = {}
OUT_NAMES = {"gemm16": "ncu_gemm_f16", "attn16": "ncu_attention_f16", "vq": "ncu_vq", "ln": "ncu_layernorm"}
METRICS += ["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]


def ncu_raw(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    return head, units, rows[2:]


def summarise_rep(tag: str, src: str, dst: str, rnd: str) -> None:
    rep = os.path.join(src, f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    head, units, rows = ncu_raw(rep)
    kcol = head.index("Kernel Name")
    seen = {}
    for r in rows:               # keep the LAST launch of each kernel (warm caches, steady clocks)
        seen[r[kcol]] = r
    import datetime
    when = datetime.datetime.fromtimestamp(os.path.getmtime(rep), datetime.timezone.utc).strftime("%Y-%m-%d %H:%MZ")
    lines = [HEADERS[tag], f"# source: ncu --set full --clock-control none, captured {when} (report under gpurun_out/, summarised by "
             "tools/summarize_profiles.py; last launch of each kernel)\n"]
    if tag in STALE_NOTES:
        lines.append(STALE_NOTES[tag])
    for name, r in seen.items():
        lines.append(f"\n## {name[:140]}\n")
        for m in METRICS:
            if m in head:
                i = head.index(m)
                lines.append(f"{m:<80s} {r[i]:>14s} {units[i]}\n")
    path = os.path.join(dst, f"{rnd}_{OUT_NAMES[tag]}.txt")
    with open(path, "w") as f:
        f.writelines(lines)
    print("wrote", path)


def summarise_launches(src: str, dst: str, rnd: str) -> None:
    path = os.path.join(src, "launches.csv")
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(dst, f"{rnd}_launches_bench_B128.csv"))
    with open(path) as f:
        text = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(text))
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    gemm = collections.defaultdict(list)
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"])
        name = r["Kernel Name"]
        short = re.sub(r"[<(].*", "", name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", ""))
        tot[short] += ns
        cnt[short] += 1
        m = re.search(r"gemm_tc_kernel<(\d+), (\d+), (\d+), (\d+), (\d+), (\d+)>", name)
        if m:
            gemm[m.groups()].append(ns / 1e3)
    total = sum(tot.values())
    bench_line = ""
    bjson = os.path.join(src, "bench.json")
    if os.path.exists(bjson):
        try:
            j = json.loads(open(bjson).read().strip().splitlines()[-1])
            bench_line = (f"# bench.py (same command, NOT under ncu): {j['value']:.1f} {j['unit']}, {j['ms_per_step']:.1f} ms/step; live CUDA-event share "
                          f"of the step spent in gemm_tc_kernel: {j.get('roofline', {}).get('share_of_step')}\n")
        except Exception:
            pass
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -s 2470 -c 810 python bench.py --steps 1 --warmup 3 --extras ''\n",
             "# (B=128/GPU, base config, fp16 data path).  ~806 launches = one fwd+bwd step; per-launch times are cold-cache and\n",
             "# serialised: compare SHARES with bench.py's live numbers, not absolutes.\n", bench_line]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{100 * v / total:6.2f}%  launches={cnt[k]:4d}  avg={v / cnt[k] / 1e3:9.1f} us  {k}\n")
    lines.append("\n# gemm_tc_kernel<KIND, BN, CG, AMAJ, BMAJ, OUT16> launches by duration (us rounded to 20: count)\n")
    for k, v in sorted(gemm.items(), key=lambda kv: -sum(kv[1])):
        hist = collections.Counter(int(round(x / 20.0) * 20) for x in v)
        lines.append(f"{k} n={len(v)} total={sum(v) / 1e3:.1f} ms  {sorted(hist.items())}\n")
    out = os.path.join(dst, f"{rnd}_launch_shares.txt")
    with open(out, "w") as f:
        f.writelines(lines)
    print("wrote", out)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r02")
    ap.add_argument("--src", default="gpurun_out/evidence")
    ap.add_argument("--dst", default="profiles")
    a = ap.parse_args()
    os.makedirs(a.dst, exist_ok=True)
    b = os.path.join(a.src, "bench.json")
    if os.path.exists(b) and open(b).read().strip().startswith("{"):
        shutil.copy(b, os.path.join(a.dst, f"{a.round}_bench_1gpu.json"))
        print("copied bench.json")
    summarise_launches(a.src, a.dst, a.round)
    for tag in ("gemm16", "attn16", "vq", "ln"):
        summarise_rep(tag, a.src, a.dst, a.round)
    return 0


if __name__ == "__main__":
    sys.exit(main())
