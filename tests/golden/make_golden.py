"""Generates tests/golden/digests.json from the C oracle (run from the repo root).

There is no reference implementation to generate vectors from; these vectors pin
the oracle against regressions and give the -m gpu tests committed targets."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle  # noqa: E402

CASES = [
    # name, dims, scene, scene_seed, seed, checkpoints
    ("config1_64_sand_block", (64, 64, 64), 1, 1, 1, [1, 2, 3, 4, 10, 50, 100, 250, 500]),
    ("mixed_64", (64, 64, 64), 2, 1, 1, [1, 2, 3, 4, 8, 64, 200]),
    ("random_96x40x24", (96, 40, 24), 3, 1, 7, [1, 2, 3, 4, 5, 6, 7, 8, 40]),
    ("mixed_noise_128x48x32", (128, 48, 32), 4, 3, 9, [1, 4, 16, 60]),
    ("odd_dims_32x7x5", (32, 7, 5), 3, 2, 3, [1, 2, 3, 4, 9]),
]

out = {"schedule_version": oracle.lib().fs3d_oracle_schedule_version(), "cases": []}
for name, dims, scene, sseed, seed, cps in CASES:
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, sseed)
    case = {"name": name, "dims": list(dims), "scene": scene, "scene_seed": sseed, "seed": seed,
            "digest0": hex(oracle.digest(g)), "digests": []}
    t = 0
    for upto in cps:
        oracle.run(g, seed, t, upto - t)
        t = upto
        case["digests"].append([upto, hex(oracle.digest(g))])
    case["histogram"] = [int(v) for v in oracle.histogram(g)[:4]]
    out["cases"].append(case)

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "digests.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote digests.json")

# Schedule version 2 (SCHEDULE.md §7): the same kind of vectors from the version-2 oracle
CASES_V2 = [
    ("random8_64", (64, 64, 64), 5, 1, 1, [1, 2, 3, 4, 8, 64, 200]),
    ("mixed8_96x64x32", (96, 64, 32), 6, 3, 9, [1, 2, 3, 4, 16, 60, 150]),
    ("random8_odd_32x7x5", (32, 7, 5), 5, 2, 3, [1, 2, 3, 4, 9]),
    ("mixed8_2048x12x6", (2048, 12, 6), 6, 1, 4, [1, 4, 13]),
]
out2 = {"schedule_version": 2, "cases": []}
for name, dims, scene, sseed, seed, cps in CASES_V2:
    nx, ny, nz = dims
    g = oracle.generate(nx, ny, nz, scene, sseed)
    case = {"name": name, "dims": list(dims), "scene": scene, "scene_seed": sseed, "seed": seed,
            "digest0": hex(oracle.digest(g)), "digests": []}
    t = 0
    for upto in cps:
        oracle.run(g, seed, t, upto - t, version=2)
        t = upto
        case["digests"].append([upto, hex(oracle.digest(g))])
    case["histogram"] = [int(v) for v in oracle.histogram(g)[:8]]
    out2["cases"].append(case)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "digests_v2.json"), "w") as f:
    json.dump(out2, f, indent=1)
print("wrote digests_v2.json")

# A golden STATE in the checkpoint format (include/fs3d.h): the MIXED_NOISE scene 64 x 32 x 16 (scene seed 6)
# after 41 steps under seed 12 — an odd step index, so a resumed run must keep the schedule phase.
from fallingsand3d_b200 import checkpoint  # noqa: E402

g = oracle.generate(64, 32, 16, 4, 6)
oracle.run(g, 12, 0, 41)
here = os.path.dirname(os.path.abspath(__file__))
checkpoint.write(os.path.join(here, "mixed_noise_64x32x16_t41.fs3d"), g, step=41, seed=12)
oracle.run(g, 12, 41, 23)
with open(os.path.join(here, "mixed_noise_64x32x16_t41.json"), "w") as f:
    json.dump({"file": "mixed_noise_64x32x16_t41.fs3d", "dims": [64, 32, 16], "scene": 4, "scene_seed": 6, "seed": 12,
               "step": 41, "resume_steps": 23, "digest_after_resume": hex(oracle.digest(g)),
               "histogram": [int(v) for v in oracle.histogram(g)[:4]]}, f, indent=1)
print("wrote mixed_noise_64x32x16_t41.fs3d")
