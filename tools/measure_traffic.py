#!/usr/bin/env python
"""DRAM traffic of the step kernel, measured with the SHIPPED library (run on the GPU box, under gpurun):

    python tools/measure_traffic.py            # plain run first, then the same command under ncu; writes
                                               # gpurun_out/step_kernel_traffic.json (+ the raw ncu csv)

bench.py's ``roofline.traffic`` / ``frac_dram`` read profiles/step_kernel_traffic.json and use it only while its
``source_fingerprint`` equals the sha256 of the kernel sources being timed (bench.source_fingerprint), so a stale
measurement can never be attached to a changed kernel.  The launches profiled are bench.py's own: 2^20 north-star
hover envs, 4 rotating state sets, random actions, Philox auto-reset (ncu additionally flushes the caches between
its replays, so the figure is the cold-cache one, like the bench's rotating sets).
"""
from __future__ import annotations

import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
WARM, COUNT = 12, 8


def driver():
    """the workload: bench.py's headline loop, WARM + COUNT launches"""
    import torch
    from uav_reinforcement_learning_control_b200 import config as Q
    from uav_reinforcement_learning_control_b200.engine import Engine
    import bench
    n = bench.NUM_ENVS
    eng = Engine(Q.EnvConfig.north_star(seed=0), n, device=0)
    dev = torch.device("cuda", 0)
    S = bench.STATE_SETS
    states = [eng.new_state() for _ in range(S)]
    obs = [torch.empty(n, 12, device=dev) for _ in range(S)]
    rew = [torch.empty(n, device=dev) for _ in range(S)]
    done = [torch.empty(n, device=dev) for _ in range(S)]
    acts = [torch.rand(n, 4, device=dev) * 2 - 1 for _ in range(S)]
    for j in range(S):
        states[j][26] = torch.full((n,), j * 1000, dtype=torch.int32, device=dev).view(torch.float32)
        eng.reset(states[j], obs=obs[j])
    for l in range(WARM + COUNT):
        j = l % S
        eng.step(states[j], acts[(j + l // S) % S], obs=obs[j], reward=rew[j], done=done[j])
    torch.cuda.synchronize()
    print("driver ok", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--driver":
        return driver()
    os.makedirs(OUT, exist_ok=True)
    cmd = [sys.executable, os.path.abspath(__file__), "--driver"]
    subprocess.check_call(cmd)                                   # must exit 0 without ncu first
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3,
             "nsecond": 1e-3}

    def capture(tag, extra):
        raw = os.path.join(OUT, f"step_kernel_traffic_ncu{tag}.csv")
        subprocess.check_call(["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
                               "--clock-control", "none"] + extra + ["-k", "regex:step2?_kernel", "-s", str(WARM), "-c", str(COUNT),
                               "--csv", "--log-file", raw] + cmd)
        rd, wr, dur = [], [], []
        with open(raw) as f:
            rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
        hdr = rows[0]
        iname, ival, iunit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rows[1:]:
            v = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1.0)
            {"dram__bytes_read.sum": rd, "dram__bytes_write.sum": wr, "gpu__time_duration.sum": dur}.get(r[iname], []).append(v)
        return rd, wr, dur

    # (1) ncu's default cache control: caches flushed before every profiled launch -> what ONE cold launch moves while it
    #     runs (dirty lines still in the 126 MB L2 when the kernel ends are NOT counted: they are written back later)
    rd, wr, dur = capture("", [])
    # (2) no cache control: every launch also absorbs the write-backs the previous launch left behind, i.e. the steady
    #     state of back-to-back launches over rotating state sets (bench.py's loop)
    rd2, wr2, dur2 = capture("_steady", ["--cache-control", "none"])
    import bench
    n = bench.NUM_ENVS
    rec = {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / max(len(rd), 1),
           "dram_read_bytes_per_launch": sum(rd) / max(len(rd), 1), "dram_write_bytes_per_launch": sum(wr) / max(len(wr), 1),
           "algorithmic_bytes_per_launch": 280 * n, "launches": len(rd), "us_per_launch_under_ncu": sum(dur) / max(len(dur), 1),
           "steady_state": {"dram_bytes_per_launch": (sum(rd2) + sum(wr2)) / max(len(rd2), 1),
                            "dram_read_bytes_per_launch": sum(rd2) / max(len(rd2), 1),
                            "dram_write_bytes_per_launch": sum(wr2) / max(len(wr2), 1),
                            "us_per_launch_under_ncu": sum(dur2) / max(len(dur2), 1),
                            "how": "same capture with --cache-control none: each launch also takes the previous launch's write-backs"},
           "source_fingerprint": bench.source_fingerprint(),
           "how": "tools/measure_traffic.py (cold launch, ncu flushes caches before each): ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on "
                  f"{COUNT} step2_kernel launches of bench.py's loop (2^20 envs, {bench.STATE_SETS} rotating state sets), shipped libquadsim.so"}
    with open(os.path.join(OUT, "step_kernel_traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
