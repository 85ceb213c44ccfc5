"""A/B harness for kernel tuning variants (run on the GPU box): for each build/variants/libquadsim_*.so
run bench.py with QS_LIB_PATH pointing at it and print one summary line."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:] or ["--steps", "100", "--warmup", "5", "--cpu-seconds", "0.5"]
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libquadsim_*.so")))
for lib in libs:
    env = dict(os.environ, QS_LIB_PATH=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    name = os.path.basename(lib)[len("libquadsim_"):-3]
    if not line:
        print(name, "FAILED", out.stderr[-400:])
        continue
    d = json.loads(line[-1])
    msg = f"{name:10s} value {d['value']:.3e}  ms/step {d['ms_per_step']:.4f}  hbm frac {d['roofline']['frac']:.3f}  e2e {d['e2e']['value']:.3e}"
    for k in ("resident", "rollout", "rollout_tc", "rollout_large", "rollout_large_tc"):
        if k in d:
            msg += f"  {k} {d[k]['value']:.3e}"
    print(msg, flush=True)
