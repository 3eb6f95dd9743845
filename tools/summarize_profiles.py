#!/usr/bin/env python
"""Turn the raw GPU evidence a `gpurun` trip leaves under gpurun_out/ into the tracked summaries under
profiles/ (the .ncu-rep files themselves are too big to commit).

    python tools/summarize_profiles.py [--round r01] [--src gpurun_out] [--dst profiles]

inputs (all optional, whatever exists is summarised):
    <src>/bench.json              bench.py line of the default run          -> <round>_bench_1gpu.json
    <src>/launches.csv            ncu --metrics gpu__time_duration.sum --csv of the same bench command
                                                                            -> <round>_launches_bench_B128.csv, <round>_launch_shares.txt
    <src>/prof_{gemm2,attn,vq,ln}.ncu-rep   ncu --set full captures of tests/ncu_target.py
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
    "gemm2": "# gemm_tf32_kernel<256,2,0,0>  to_qkv forward shape M=131072 N=2304 K=768 (round_out=1)\n"
             "# algorithmic: 463.9 GFLOP, 1.618 GB (A 403 MB + B 7 MB + C 1208 MB).  traffic = dram read+write below.  python tests/ncu_target.py gemm2\n",
    "attn": "# attention kernels, B=32 N=1024 heads=12 dh=64 (python tests/ncu_target.py attn)\n"
            "# algorithmic flops: fwd 4*B*H*N*N*dh = 103 GFLOP; bwd 10*B*H*N*N*dh = 258 GFLOP (dKV 6, dQ 4 GEMM-units of 2*N*N*dh)\n",
    "vq": "# vq_fwd_kernel 131072 tokens x 8192 codes x 32 dims (python tests/ncu_target.py vq)\n"
          "# algorithmic: 68.7 GFLOP fp32 FMA, 34.6 MB (z in, z_q out, idx out, codebook)\n",
    "ln": "# ln_fwd / ln_bwd kernels M=131072 D=768 (python tests/ncu_target.py ln)\n"
          "# algorithmic bytes: fwd 805 MB (read x, write y), bwd 1611 MB (read dy, x, dres; write dx)\n",
}
# captures that predate later kernel changes (the round's GPU budget ran out before they could be redone)
STALE_NOTES = {
    "attn": "# NOTE: this capture predates the TMA-store epilogues and the pipelined forward softmax (commit e159a95 and later).\n"
            "#       Current per-launch times at B=128 from the round-end launch list (r01_launch_shares.txt): fwd 906 us, dKV 1505 us,\n"
            "#       dQ 1261 us (this capture's build: 1223 / 2145 / 1387 us); at B=64 under CUDA events: fwd 0.445 ms, bwd 1.421 ms.\n",
    "ln": "# NOTE: captured after the shared-memory-accumulator rewrite of ln_bwd but before the optional third accumulator\n"
          "#       (colsum of dx); round-end launch list: ln_bwd 256 us in the training step (NACC=3 variant), ln_fwd 126 us.\n",
}
OUT_NAMES = {"gemm2": "ncu_gemm_cg2", "attn": "ncu_attention", "vq": "ncu_vq", "ln": "ncu_layernorm"}


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
        short = re.sub(r"<.*", "", name.replace("void ", ""))
        tot[short] += ns
        cnt[short] += 1
        m = re.search(r"gemm_tf32_kernel<(\d+), (\d+), (\d+), (\d+)>", name)
        if m:
            gemm[m.groups()].append(ns / 1e3)
    total = sum(tot.values())
    bench_line = ""
    bjson = os.path.join(src, "bench.json")
    if os.path.exists(bjson):
        try:
            j = json.loads(open(bjson).read().strip().splitlines()[-1])
            bench_line = (f"# bench.py (same command, NOT under ncu): {j['value']:.1f} {j['unit']}, {j['ms_per_step']:.1f} ms/step; live CUDA-event share "
                          f"of the step spent in gemm_tf32_kernel: {j.get('roofline', {}).get('share_of_step')}\n")
        except Exception:
            pass
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -s 2800 -c 1000 python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n",
             "# (B=128/GPU, base config).  ~1000 launches = a little over one fwd+bwd step; per-launch times are cold-cache and\n",
             "# serialised: compare SHARES with bench.py's live numbers, not absolutes.\n", bench_line]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{100 * v / total:6.2f}%  launches={cnt[k]:4d}  avg={v / cnt[k] / 1e3:9.1f} us  {k}\n")
    lines.append("\n# gemm_tf32_kernel<BN, CG, AMAJ, BMAJ> launches by duration (us rounded to 20: count)\n")
    for k, v in sorted(gemm.items(), key=lambda kv: -sum(kv[1])):
        hist = collections.Counter(int(round(x / 20.0) * 20) for x in v)
        lines.append(f"{k} n={len(v)} total={sum(v) / 1e3:.1f} ms  {sorted(hist.items())}\n")
    out = os.path.join(dst, f"{rnd}_launch_shares.txt")
    with open(out, "w") as f:
        f.writelines(lines)
    print("wrote", out)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r01")
    ap.add_argument("--src", default="gpurun_out")
    ap.add_argument("--dst", default="profiles")
    a = ap.parse_args()
    os.makedirs(a.dst, exist_ok=True)
    b = os.path.join(a.src, "bench.json")
    if os.path.exists(b) and open(b).read().strip().startswith("{"):
        shutil.copy(b, os.path.join(a.dst, f"{a.round}_bench_1gpu.json"))
        print("copied bench.json")
    summarise_launches(a.src, a.dst, a.round)
    for tag in ("gemm2", "attn", "vq", "ln"):
        summarise_rep(tag, a.src, a.dst, a.round)
    return 0


if __name__ == "__main__":
    sys.exit(main())
