#!/usr/bin/env python
"""A/B helper for kernel tuning on a GPU box: run bench.py (device-resident leg only) once per library variant and print
one summary line each -- step ms, per-kernel ms, whether the step still reproduces the reference's canvas (verify).
    python profiles/ab_bench.py default u1 u4 ...      (variants built by profiles/build_variant.sh; "default" = product lib)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "super-resolution-system_b200")
extra = [a for a in sys.argv[1:] if a.startswith("--")]
for tag in [a for a in sys.argv[1:] if not a.startswith("--")]:
    env = dict(os.environ)
    if tag != "default":
        env["SRF_LIB"] = os.path.join(PKG, f"libsrfuse_{tag}.so")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e", "--no-cpu-baseline"] + extra,
                         capture_output=True, text=True, env=env, cwd=ROOT)
    if res.returncode != 0 or not res.stdout.strip():
        print(tag, "FAILED", res.returncode, res.stderr[-1500:])
        continue
    d = json.loads(res.stdout.strip().splitlines()[-1])
    ks = " ".join(f"{k['kernel']}={k['ms_per_step']:.3f}" for k in d["kernels"])
    v = d.get("verify") or {}
    print(f"{tag}: step {d['ms_per_step']:.3f} ms | {ks} | verify ok={v.get('ok')} sha={v.get('canvas_sha256_match')}", flush=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ab_{tag}.json"), "w") as f:
        json.dump(d, f)
