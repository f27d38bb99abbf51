"""Generates tests/golden/bench_digests.json: oracle digests of bench.py's workload (the 2048^3 MIXED_NOISE scene,
scene seed 1, coin seed 1 — BASELINE configs[3]) after the step counts bench.py reaches, computed by the CPU oracle
on the FULL grid (8 GiB; ~10 s per step on 8 cores).  bench.py asserts its own digest against these, and
tests/test_step_gpu.py::test_headline_grid_matches_oracle_digests replays them on the GPU.

Run from the repo root:  python tests/golden/make_bench_digests.py [last_step]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle  # noqa: E402

N = 2048
LAST = int(sys.argv[1]) if len(sys.argv) > 1 else 106
# steps bench.py can end on: warm-up rounded up to even (4 or 6 for --warmup 3..6) + K for K = 20, 50, 100, and a few early ones
# (warm-up rounded up to a multiple of four — so that every timed pass of a single-GPU run is a four-step pass — + K)
KEEP = sorted(set([1, 2, 3, 4, 6, 8, 13, 24, 26, 28, 54, 56, 58, 104, 106, 108]))
out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_digests.json")

g = oracle.generate(N, N, N, 4, 1)
old = {}
if os.path.exists(out_path):                      # keep what an earlier (longer) run already established
    with open(out_path) as f:
        old = json.load(f).get("digests", {})
out = {"schedule_version": oracle.lib().fs3d_oracle_schedule_version(), "dims": [N, N, N], "scene": 4, "scene_seed": 1,
       "seed": 1, "digest0": hex(oracle.digest(g)), "histogram": [int(v) for v in oracle.histogram(g)[:4]], "digests": {}}
# a thin slab's step-0 digest, cheap to regenerate: lets the CPU test suite pin the scene the big run started from
out["slab_1020_1028_digest0"] = hex(oracle.digest(oracle.generate(N, N, N, 4, 1, 1020, 1028), 1020))
t0 = time.time()
for t in range(LAST):
    oracle.step(g, 1, t)
    if t + 1 in KEEP:
        out["digests"][str(t + 1)] = hex(oracle.digest(g))
        assert old.get(str(t + 1), out["digests"][str(t + 1)]) == out["digests"][str(t + 1)], "oracle digest changed"
        for k, v in old.items():
            out["digests"].setdefault(k, v)
        with open(out_path, "w") as f:
            json.dump(out, f, indent=1)
        print(f"step {t + 1}: {out['digests'][str(t + 1)]}  ({time.time() - t0:.0f} s)", flush=True)
print("wrote", out_path)
