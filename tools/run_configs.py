#!/usr/bin/env python
"""Runs the BASELINE.json configs and writes one JSON report (profiles/r01_configs.json).

  python tools/run_configs.py 1 2 3            (single GPU)
  torchrun --nproc-per-node 8 tools/run_configs.py 5      (config 5: 4096^3 on 8 GPUs + raymarch)

Config 4 (2048^3 strong scaling) is bench.py itself.  The CPU numbers are the builder-written oracle
(there is no reference CPU update: SURVEY.md §0), timed on this box's host cores.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import fallingsand3d_b200 as fs3d  # noqa: E402


def cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def config1():
    """64^3, single sand block, 500 steps: GPU digest == oracle digest after every step."""
    from oracle import oracle
    n = 64
    g = oracle.generate(n, n, n, 1, 1)
    t0 = time.perf_counter()
    dig = []
    for t in range(500):
        oracle.step(g, 1, t)
        dig.append(oracle.digest(g))
    cpu_s = time.perf_counter() - t0
    with fs3d.VoxelWorld(n, n, n, seed=1) as w:
        w.generate(fs3d.SCENE_SAND_BLOCK, 1)
        ok = True
        for t in range(500):
            w.step(1)
            ok = ok and w.digest() == dig[t]
        h = w.histogram()
        w2 = fs3d.VoxelWorld(n, n, n, seed=1)
        w2.generate(fs3d.SCENE_SAND_BLOCK, 1)
        ms, _ = w2.step_timed(500)
        w2.close()
    return {"config": "64^3 sand block, 500 steps", "bit_exact_every_step": bool(ok),
            "histogram": {"EMPTY": int(h[0]), "SAND": int(h[1])},
            "gpu_ms_total": ms, "gpu_voxel_updates_per_s": n ** 3 * 500 / (ms * 1e-3),
            "cpu_oracle_s_total": cpu_s, "cpu_oracle_voxel_updates_per_s": n ** 3 * 500 / cpu_s, "cpu_cores": cores(),
            "note": "CPU = builder-written oracle run headless (includes a digest per step); not a reference number"}


def config2():
    """256^3 mixed, 1000 steps: digest every 10 steps + full compare at the end."""
    from oracle import oracle
    n = 256
    g = oracle.generate(n, n, n, 2, 1)
    with fs3d.VoxelWorld(n, n, n, seed=1) as w:
        w.generate(fs3d.SCENE_MIXED, 1)
        h0 = w.histogram()
        ok = w.digest() == oracle.digest(g)
        cpu_s = 0.0
        for t in range(0, 1000, 10):
            w.step(10)
            t0 = time.perf_counter()
            oracle.run(g, 1, t, 10)
            cpu_s += time.perf_counter() - t0
            ok = ok and w.digest() == oracle.digest(g)
        full = bool(np.array_equal(w.download(), g))
        hist_ok = bool(np.array_equal(w.histogram(), h0))
        w2 = fs3d.VoxelWorld(n, n, n, seed=1)
        w2.generate(fs3d.SCENE_MIXED, 1)
        ms, _ = w2.step_timed(1000)
        w2.close()
    return {"config": "256^3 mixed sand + water + stone, 1000 steps", "digest_equal_every_10_steps": bool(ok),
            "full_grid_equal_at_end": full, "histogram_invariant": hist_ok,
            "gpu_ms_total": ms, "gpu_voxel_updates_per_s": n ** 3 * 1000 / (ms * 1e-3),
            "cpu_oracle_s_total": cpu_s, "cpu_oracle_voxel_updates_per_s": n ** 3 * 1000 / cpu_s, "cpu_cores": cores()}


def config3():
    """1024^3 random 50 % fill: sustained throughput, skipping off and on."""
    n = 1024
    out = {"config": "1024^3 random fill (25 % sand, 25 % water), sustained"}
    for name, flags in (("skip_off_fused", 0), ("skip_off_single_step", fs3d.FLAG_NO_FUSE)):
        with fs3d.VoxelWorld(n, n, n, seed=1, flags=flags) as w:
            w.generate(fs3d.SCENE_RANDOM, 1)
            h0 = w.histogram()
            w.step(20)
            ms, _ = w.step_timed(200)
            assert np.array_equal(w.histogram(), h0)
            out[name] = {"ms_per_step": ms / 200, "voxel_updates_per_s": n ** 3 * 200 / (ms * 1e-3),
                         "roofline_frac_2B_per_update": 2 * n ** 3 * 200 / (ms * 1e-3) / 6545e9}
    with fs3d.VoxelWorld(n, n, n, seed=1, flags=fs3d.FLAG_SKIP_SETTLED) as w:
        w.generate(fs3d.SCENE_RANDOM, 1)
        h0 = w.histogram()
        trace = []
        for k in range(40):                       # 4000 steps: the column settles, the water surface never does
            ms, _ = w.step_timed(100)
            run, total = w.activity()
            trace.append({"steps": (k + 1) * 100, "ms_per_step": ms / 100, "active_tile_fraction": run / total,
                          "nominal_voxel_updates_per_s": n ** 3 * 100 / (ms * 1e-3)})
        assert np.array_equal(w.histogram(), h0)
        out["skip_on"] = {"trace": trace, "digest": hex(w.digest()),
                          "note": "nominal rate counts skipped voxels; do not read a roofline % from it"}
    with fs3d.VoxelWorld(n, n, n, seed=1) as w:    # same run without skipping must land on the same state
        w.generate(fs3d.SCENE_RANDOM, 1)
        w.step(4000)
        out["skip_on"]["digest_equals_skip_off"] = hex(w.digest()) == out["skip_on"]["digest"]
    return out


def config5():
    """4096^3 (68.7 G voxels) on 8 GPUs, raymarch 1920x1080 every 10 steps."""
    import torch
    import torch.distributed as dist
    from fallingsand3d_b200.slab import SlabWorld
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(os.environ.get("FS3D_CONFIG5_SIZE", "4096"))
    sw = SlabWorld(n, n, n, seed=1)
    sw.generate(fs3d.SCENE_RANDOM, 1)
    h0 = sw.histogram()
    sw.step(4)
    sw.sync()
    dist.barrier()
    cam = dict(pos=(0.0, 0.0, -1.6), yaw_deg=0.0, aspect=1920.0 / 1080.0, width=1920, height=1080, mode=fs3d.RM_VOXELS)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (a) stepping alone
    ev0.record(sw.engine.stream)
    sw.step(60)
    ev1.record(sw.engine.stream)
    sw.sync()
    t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / 60
    # (b) 60 steps with a frame every 10 steps, wall clock
    dist.barrier()
    t0 = time.perf_counter()
    rm = []
    phases = []
    img = None
    for k in range(6):
        sw.step(10)
        sw.sync()                          # the frame timer must not include the ten steps still in flight
        t1 = time.perf_counter()
        img = sw.raymarch(**cam)
        rm.append(time.perf_counter() - t1)
        phases.append(dict(getattr(sw, "last_frame_ms", {})))
    sw.sync()
    dist.barrier()
    wall = time.perf_counter() - t0
    ok = bool(np.array_equal(sw.histogram(), h0))
    res = None
    if rank == 0:
        res = {"config": f"{n}^3 random fill on {sw.world_size} GPUs, raymarch 1920x1080 every 10 steps",
               "ms_per_step_stepping_only": ms_step, "voxel_updates_per_s": n ** 3 / (ms_step * 1e-3),
               "wall_s_60_steps_with_6_frames": wall, "raymarch_ms_per_frame_incl_gather": [1e3 * r for r in rm], "raymarch_phases_ms_rank0": phases,
               "image_nonblack_pixels": int((img[..., :3].sum(axis=-1) > 0).sum()), "histogram_invariant": ok,
               "digest": hex(sw.digest())}
    else:
        sw.digest()
    sw.close()
    dist.barrier()
    dist.destroy_process_group()
    return res


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3"]
    report = {}
    for c in which:
        r = {"1": config1, "2": config2, "3": config3, "5": config5}[c]()
        if r is not None:
            report[f"config{c}"] = r
            print(json.dumps({f"config{c}": r}), flush=True)
    if report:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "configs_" + "_".join(which) + ".json"), "w") as f:
            json.dump(report, f, indent=1)
