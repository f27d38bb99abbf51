#!/usr/bin/env python
"""bench.py — voxel-updates/s of the fs3d step on N B200s, as one JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl ours|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, NCCL halo exchange)

Workload (BASELINE.json configs[3], the one the 60 %-of-roofline target is quoted on): a 2048³
MIXED_NOISE scene (stone floor + obstacles + sand box + water box + sand/water noise in the upper
half), generated on device, strong-scaled as z-slabs over the ranks.  One step = one full
SCHEDULE.md step of every voxel.  The grid (8 GiB per buffer) is far larger than the 126 MB L2, so
no L2 flush is needed between timed steps.

  value      device-resident throughput: nx·ny·nz·K / (max over ranks of the CUDA-event time of K steps)
  e2e        same metric through the C ABI with HOST buffers: every step uploads the grid from pinned
             host memory, steps, and downloads the result (PCIe inside the timed region)
  roofline   2 B per voxel-update (1 B read + 1 B written) over the measured HBM copy peak
  cpu_baseline  the CPU oracle (a builder-written port: the reference has NO implementation of this
             path) timed on this box's host cores on a bounded sample

--impl reference times that same CPU oracle on all host threads (rank 0 only) on a bounded sample
of the same workload: it is the only "reference CPU implementation" that can exist.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel-updates/s"
SCENE_MIXED_NOISE = 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def wait_ready(self, timeout=5.0):
        """nvidia-smi takes a while to start: block until its first sample has arrived."""
        t0 = time.time()
        while self.proc and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """Call right before the timed region: only later samples are reported."""
        self.first = len(self.samples)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for s in self.samples[getattr(self, "first", 0):]:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    """(threads the oracle's OpenMP loops really use, OMP_NUM_THREADS as seen, cpus in the affinity mask)."""
    from oracle import oracle
    try:
        aff = len(os.sched_getaffinity(0))
    except Exception:
        aff = os.cpu_count() or 1
    return oracle.threads(), os.environ.get("OMP_NUM_THREADS"), aff


def sample_planes(n, threads, budget_voxels):
    """z-planes of the CPU sample: about `budget_voxels`, but never fewer than 8 planes per host thread — the oracle's
    work items are (plane, block-row) pairs and (plane-pair, block-row) pairs, so even the thinnest sample gives every
    thread thousands of items (round 1 sampled 16 planes for 16-32 threads and parallelised over z only)."""
    nz = max(budget_voxels // (n * n), 8 * threads, 2)
    nz = min(n, nz + (nz & 1))
    return int(nz)


def cpu_oracle_rate(nx, ny, nz_sample, steps, seed=1):
    """voxel-updates/s of the CPU oracle on a bounded slab sample of the workload."""
    from oracle import oracle
    import numpy as np
    full_nz = nx  # cubic workload
    zlo = full_nz // 2 - nz_sample // 2      # straddles the noisy upper half and the obstacles
    g = oracle.generate(nx, ny, full_nz, SCENE_MIXED_NOISE, 1, zlo, zlo + nz_sample)
    oracle.step(g, seed, 0)                   # warm-up (page faults, OpenMP pool)
    t0 = time.perf_counter()
    for t in range(1, 1 + steps):
        oracle.step(g, seed, t)
    dt = time.perf_counter() - t0
    cores = oracle.threads()
    return nx * ny * nz_sample * steps / dt, cores, dt, f"{nx}x{ny}x{nz_sample} slab of the {nx}^3 MIXED_NOISE scene, {steps} steps"


def reference_arm(args):
    """--impl reference: the CPU oracle (kind "port") on all host threads, bounded sample.

    Each step is one oracle step of a fixed slab sample of the workload (~64 Mi voxels), so the whole
    --steps K --warmup W run ends within a few minutes whatever K is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm runs alone on rank 0 and is meant to use every
    # host thread it can, so undo that before libgomp is loaded (the oracle library is the first OpenMP user here)
    if "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ.pop("OMP_NUM_THREADS", None)
    from oracle import oracle
    n = args.size
    threads, omp_env, affinity = host_threads()
    nz_sample = sample_planes(n, threads, 1 << 28)      # the same sample the cpu_baseline leg of the GPU arm times
    zlo = n // 2 - nz_sample // 2
    g = oracle.generate(n, n, n, SCENE_MIXED_NOISE, 1, zlo, zlo + nz_sample)
    sample = f"{n}x{n}x{nz_sample} slab (z from {zlo}) of the {n}^3 MIXED_NOISE scene, one oracle step per bench step"
    t = 0
    for _ in range(args.warmup):
        oracle.step(g, 1, t); t += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.step(g, 1, t); t += 1
    dt = time.perf_counter() - t0
    cores = threads
    value = n * n * nz_sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxel-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{n}^3 MIXED_NOISE scene (BASELINE configs[3]); CPU oracle on a bounded sample",
                   "grid": [n, n, n], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "voxel-updates/s", "cores": cores, "kind": "port", "sample": sample,
                         "omp_num_threads_env": omp_env, "affinity_cpus": affinity,
                         "note": "the reference has no CPU update to time (SURVEY.md §0); this is the builder-written "
                                 "CPU oracle of the same schedule, OpenMP over (plane, block-row) work items"},
        "e2e": {"value": value, "unit": "voxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class numa_local:
    """Context manager: binds this thread to the CPUs of the NUMA node GPU `local_rank` hangs off (sysfs
    local_cpulist of its PCI function) while pinned host memory is allocated and first touched, then restores the
    affinity mask (so the CPU baseline afterwards still sees every core).  Yields a small report dict."""

    def __init__(self, local_rank):
        self.local_rank = local_rank
        self.saved = None
        self.report = {"bound": False}

    def __enter__(self):
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.local_rank)
            bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            base = os.path.join("/sys/bus/pci/devices", bdf)
            with open(os.path.join(base, "local_cpulist")) as f:
                spec = f.read().strip()
            cpus = set()
            for part in spec.split(","):
                if "-" in part:
                    lo, hi = part.split("-")
                    cpus.update(range(int(lo), int(hi) + 1))
                elif part:
                    cpus.add(int(part))
            node = None
            try:
                with open(os.path.join(base, "numa_node")) as f:
                    node = int(f.read().strip())
            except Exception:
                pass
            self.saved = os.sched_getaffinity(0)
            cpus &= self.saved
            if cpus:
                os.sched_setaffinity(0, cpus)
                self.report = {"bound": True, "pci": bdf, "numa_node": node, "cpus": len(cpus)}
        except Exception as e:                     # no sysfs / no permission: allocate wherever the thread runs
            self.report = {"bound": False, "why": repr(e)[:120]}
        return self.report

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def golden_digest_check(n, total_steps, digest):
    """Compares the run's digest with the CPU oracle's digest of the same workload at the same step count
    (tests/golden/bench_digests.json, generated by the oracle on the FULL grid).  Raises on a mismatch."""
    p = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
    if not os.path.exists(p):
        return "no golden file"
    with open(p) as f:
        gold = json.load(f)
    if gold["dims"] != [n, n, n]:
        return f"no oracle digest for {n}^3 (golden file holds {gold['dims'][0]}^3)"
    want = gold["digests"].get(str(total_steps))
    if want is None:
        return f"no oracle digest for step {total_steps} (have {sorted(int(k) for k in gold['digests'])})"
    assert int(want, 16) == digest, f"digest after {total_steps} steps {hex(digest)} != CPU oracle's {want}"
    return f"equals the CPU oracle's digest of the full {n}^3 grid after {total_steps} steps"


def big_grid_block(args, world_size, rank, hbm_peak):
    """BASELINE configs[4]: the 4096^3 RANDOM scene (68.7 G voxels, 2 x 64 GiB) — on ONE 180 GB B200 at N = 1, as z-slabs
    at N > 1 — so that the driver's records hold T_N at 4096^3 for every N it runs (efficiency = T_1 / (N T_N)), plus
    one 1920x1080 ray-marched frame.  Same timing rules as the main measurement."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import fallingsand3d_b200 as fs3d
    from fallingsand3d_b200.slab import SlabWorld, slab_bounds
    n, K = args.big_size, args.big_steps
    voxels = n * n * n
    planes = n if world_size == 1 else max(b - a for a, b in slab_bounds(n, world_size))
    need = 2 * (planes + 2) * n * n + (1 << 30)
    free, _total = torch.cuda.mem_get_info()
    ok = torch.tensor([1 if free >= need else 0], device="cuda")
    if world_size > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        return {"skipped": f"needs {need / 2**30:.0f} GiB per GPU, {free / 2**30:.0f} GiB free"}
    try:
        if world_size == 1:
            w = fs3d.VoxelWorld(n, n, n, seed=1)
            w.generate(fs3d.SCENE_RANDOM, 1)
            h0 = w.histogram()
            w.step(4)
            w.sync()
            ms, launches = w.step_timed(K)
            w.raymarch(width=1920, height=1080, mode=fs3d.RM_VOXELS)          # first frame: allocations, module load
            t0 = time.perf_counter()
            w.raymarch(width=1920, height=1080, mode=fs3d.RM_VOXELS)
            frame_ms = (time.perf_counter() - t0) * 1e3
            assert np.array_equal(w.histogram(), h0), "material counts changed: invalid run"
            digest = w.digest()
            w.close()
        else:
            sw = SlabWorld(n, n, n, seed=1)
            sw.generate(fs3d.SCENE_RANDOM, 1)
            h0 = sw.histogram()
            sw.step(4)
            sw.sync()
            dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st = sw.engine.stream
            ev0.record(st)
            sw.step(K)
            ev1.record(st)
            sw.sync()
            torch.cuda.synchronize()
            t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
            dist.barrier()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            sw.raymarch(width=1920, height=1080, mode=fs3d.RM_VOXELS)        # first frame sets the shared frame up
            dist.barrier()
            t0 = time.perf_counter()
            sw.raymarch(width=1920, height=1080, mode=fs3d.RM_VOXELS)
            frame_ms = (time.perf_counter() - t0) * 1e3
            assert np.array_equal(sw.histogram(), h0), "material counts changed: invalid run"
            digest = sw.digest()
            launches = sw.engine.world.kernel_launches if sw.p2p else 3 * ((K + 1) // 2)
            sw.close()
    except Exception as e:                                  # the main line must still be printed
        return {"failed": repr(e)[:200]}
    achieved = 2.0 * voxels * K / (ms * 1e-3) / 1e9 / world_size
    return {"grid": [n, n, n], "scene": "RANDOM (25 % sand, 25 % water; BASELINE configs[4], SURVEY.md §8d config 5)",
            "steps": K, "warmup": 4, "ms_per_step": ms / K, "value": voxels * K / (ms * 1e-3), "unit": "voxel-updates/s",
            "n_gpus": world_size, "gpu_launches": int(launches), "digest": hex(digest), "digest_step": 4 + K,
            "roofline_frac_algorithmic": achieved / hbm_peak,
            "raymarch_1920x1080_ms": frame_ms,
            "note": "efficiency at N GPUs = (ms_per_step at N = 1) / (N x ms_per_step at N), both from this block; the digest "
                    "is the same at every N"}


def materials8_block(args, hbm_peak):
    """Schedule version 2 (FS3D_FLAG_MATERIALS8: eight materials on three bit-planes, SCHEDULE.md §7) on the same grid size:
    the MIXED8 scene, fused and unfused, beside version 1's numbers — same bytes per voxel, ~1.5x the integer work."""
    import numpy as np
    import fallingsand3d_b200 as fs3d
    n, K = args.size, args.m8_steps
    voxels = n * n * n
    out = {"grid": [n, n, n], "scene": "MIXED8 (MIXED + gas / oil / honey / gravel boxes + RANDOM8 noise in the upper half)",
           "schedule_version": 2, "steps": K, "warmup": 4}
    try:
        for name, flags in (("fused", 0), ("single_step", fs3d.FLAG_NO_FUSE)):
            with fs3d.VoxelWorld(n, n, n, seed=1, flags=fs3d.FLAG_MATERIALS8 | flags) as w:
                w.generate(fs3d.SCENE_MIXED8, 1)
                h0 = w.histogram()
                w.step(4)
                w.sync()
                ms, launches = w.step_timed(K)
                assert np.array_equal(w.histogram(), h0), "material counts changed: invalid run"
                a = 2.0 * voxels * K / (ms * 1e-3) / 1e9
                out[name] = {"ms_per_step": ms / K, "value": voxels * K / (ms * 1e-3), "gpu_launches": int(launches),
                             "achieved": a, "frac": a / hbm_peak, "digest": hex(w.digest())}
        assert out["fused"]["digest"] == out["single_step"]["digest"], "fused and unfused version-2 runs disagree"
    except Exception as e:
        out["failed"] = repr(e)[:200]
    return out


def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import fallingsand3d_b200 as fs3d
    from fallingsand3d_b200 import build as fsbuild, _lib
    from fallingsand3d_b200.slab import SlabWorld

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libfs3d has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if rank == 0:
        fsbuild.build()
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    _lib.load()

    n = args.size
    K, Wm = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = peaks()
    voxels = n * n * n
    Wwarm = (Wm + 3) // 4 * 4       # warm-up steps actually run: rounded up to a multiple of four so that the timed
                                    # steps start where a four-step pass can (and the digest is comparable across N)

    single = None
    if world_size == 1:
        # (a) one kernel pass per step — the kernel the 2 B/voxel-update roofline describes
        if not args.fused_only:
            w1 = fs3d.VoxelWorld(n, n, n, seed=1, flags=fs3d.FLAG_NO_FUSE)
            w1.generate(SCENE_MIXED_NOISE, 1)
            w1.step(Wwarm)             # the same warm-up as the fused world: digests are always comparable
            w1.sync()
            ms1, l1 = w1.step_timed(K)
            d1 = w1.digest()
            w1.close()
            a1 = 2.0 * voxels * K / (ms1 * 1e-3) / 1e9
            single = {"value": voxels * K / (ms1 * 1e-3), "ms_per_step": ms1 / K, "gpu_launches": int(l1),
                      "achieved": a1, "frac": a1 / hbm_peak, "digest": hex(d1),
                      "note": "FS3D_FLAG_NO_FUSE: one pass (1 B read + 1 B written per voxel) per step"}
        # (a2) two steps per pass (FS3D_FLAG_NO_FUSE4): the product path of every world the four-step kernel does not serve
        two = None
        if not args.fused_only:
            w2 = fs3d.VoxelWorld(n, n, n, seed=1, flags=fs3d.FLAG_NO_FUSE4)
            w2.generate(SCENE_MIXED_NOISE, 1)
            w2.step(Wwarm)
            w2.sync()
            ms2, l2 = w2.step_timed(K)
            d2 = w2.digest()
            w2.close()
            a2 = 2.0 * voxels * K / (ms2 * 1e-3) / 1e9
            two = {"value": voxels * K / (ms2 * 1e-3), "ms_per_step": ms2 / K, "gpu_launches": int(l2), "achieved": a2,
                   "frac": a2 / hbm_peak, "digest": hex(d2), "note": "FS3D_FLAG_NO_FUSE4: steps 2k, 2k+1 share a pass"}
        # (b) the product path: fs3d_step fuses four steps into one pass where it can (rows of 1024 / 2048 voxels, single
        # GPU: 0.5 B per voxel-update), else two
        w = fs3d.VoxelWorld(n, n, n, seed=1)
        w.generate(SCENE_MIXED_NOISE, 1)
        h0 = w.histogram()
        sampler = ClockSampler(local_rank)
        sampler.start()
        sampler.wait_ready()
        w.step(Wwarm)                  # a multiple of four: every timed pass is a whole fused pass
        w.sync()
        sampler.mark()
        ms, launches = w.step_timed(K)
        clocks = sampler.stop()
        assert np.array_equal(w.histogram(), h0), "material counts changed: invalid run"
        digest = w.digest()
        if single is not None:
            assert single["digest"] == hex(digest), "fused and unfused runs disagree"
            assert two["digest"] == hex(digest), "four-step and two-step passes disagree"
    else:
        sw = SlabWorld(n, n, n, seed=1)
        sw.generate(SCENE_MIXED_NOISE, 1)
        h0 = sw.histogram()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
            sampler.wait_ready()
        sw.step(Wwarm)                  # same warm-up as at N = 1, so the digest is comparable across N
        sw.sync()
        if sw.p2p:
            sw.engine.world.push_wait_stats()       # reset: only the timed region's halo waits are reported
        dist.barrier()
        torch.cuda.synchronize()
        sampler.mark()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = sw.engine.stream
        l0 = sw.engine.world.kernel_launches
        ev0.record(st)
        sw.step(K)
        ev1.record(st)
        sw.sync()
        torch.cuda.synchronize()
        launches = sw.engine.world.kernel_launches - l0 if sw.p2p else 3 * ((K + 1) // 2)     # this rank's kernels (NCCL's own not counted)
        t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        clocks = sampler.stop() if rank == 0 else None
        halo_wait = None
        if sw.p2p:
            # what inter-GPU skew cost inside the timed region: time PUSH warps spent blocked on a neighbour's counter
            ws = torch.tensor(sw.engine.world.push_wait_stats(), device="cuda", dtype=torch.float64)
            wmax = ws.clone()
            dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ws, op=dist.ReduceOp.SUM)
            fuse4 = bool(getattr(sw, "fuse4", False))
            passes = (K + 3) // 4 if fuse4 else (K + 1) // 2
            npairs = (sw.z_end - sw.z_begin) // 2 + 1
            if fuse4:       # bands of 4 pairs, one segment per unit (12 warps per SM, n / 1024 warps per unit), 7 warm-up iterations
                segs, work = 148 * 12 // max(1, n // 1024), -(-npairs // 4) * (n // 2 + 4)
                lead = min(1.0, 7.0 * segs / work)
            else:
                lead = min(1.0, 3.0 * 148 * 8 / (npairs * (n // 2 + 2)))
            halo_wait = {"longest_single_wait_ms": float(wmax[1].item()) / 1e6, "blocking_waits_all_ranks": int(ws[2].item()),
                         "blocked_warp_ms_per_pass_worst_rank": float(wmax[0].item()) / 1e6 / passes,
                         "steps_per_pass": 4 if fuse4 else 2, "lead_in_share_of_iterations": lead,
                         "note": "warps of a slab's edge pairs / bands wait (bounded) for the neighbour's previous pass; lead-in = "
                                 "re-computed iterations per march segment (3 with two steps per pass, 7 with four)"}
        assert np.array_equal(sw.histogram(), h0), "material counts changed: invalid run"
        digest = sw.digest()

    value = voxels * K / (ms * 1e-3)
    achieved = 2.0 * voxels * K / (ms * 1e-3) / 1e9 / world_size     # per-GPU algorithmic GB/s
    total_steps = Wwarm + K
    digest_check = golden_digest_check(n, total_steps, digest)
    p2p = True if world_size == 1 else bool(sw.p2p)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    # The call a host-side user makes for a grid that lives in host memory: fs3d_step_host (a rank: fs3d_slab_step_host)
    # on a pinned buffer.  The product's unit of work is the fused pass, so the headline e2e form is TWO steps per call
    # (the grid crosses PCIe once in and once out per two steps); the one-step-per-call form is reported beside it.
    e2e = None
    if args.e2e_steps <= 0:
        if world_size == 1:
            w.close()
        else:
            sw.close()
    else:
        ke = max(1, min(args.e2e_steps, K))
        with numa_local(local_rank) as numa:         # pinned pages on the GPU's own NUMA node; affinity restored after
            if world_size == 1:
                host = torch.empty((n, n, n), dtype=torch.uint8, pin_memory=True)
                hv = host.numpy()
                w.download(hv)
            else:
                host = torch.empty((sw.z_end - sw.z_begin, n, n), dtype=torch.uint8, pin_memory=True)
                hv = host.numpy()
                hv[...] = sw.download()
        stepper = w if world_size == 1 else sw
        idx = (lambda: w.step_index) if world_size == 1 else (lambda: sw.step_index)

        def timed(ns):
            while idx() % ns:
                stepper.step_host(hv, hv, 1)             # a fused pass starts on a multiple of its length
            stepper.step_host(hv, hv, ns)                # warm-up of the copy path
            if world_size > 1:
                torch.cuda.synchronize()
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                stepper.step_host(hv, hv, ns)            # returns when the stepped grid is back on the host
            if world_size > 1:
                torch.cuda.synchronize()
                dist.barrier()
            dt = time.perf_counter() - t0
            if world_size > 1:
                tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            return dt

        # the call is PCIe-bound, so what it costs per step is set by how many steps one call advances: four where the
        # four-step kernel exists (one process, rows of 1024 / 2048 / 4096 voxels), two otherwise; the shorter forms beside it
        spc = 4 if (world_size == 1 and n in (1024, 2048, 4096)) else 2
        dt4 = timed(4) if spc == 4 else None
        dt2 = timed(2)
        dt1 = timed(1)
        dtm = dt4 if spc == 4 else dt2
        how = ("fs3d_step_host(pinned host grid)" if world_size == 1 else
               "per rank: fs3d_slab_step_host on its pinned host slab (edge planes pushed to the neighbours over peer "
               "memory first, one barrier per call)")
        e2e = {"value": voxels * spc * ke / dtm, "unit": "voxel-updates/s",
               "h2d_bytes_per_step": voxels // spc, "d2h_bytes_per_step": voxels // spc,
               "steps": spc * ke, "calls": ke, "steps_per_call": spc, "ms_per_step": dtm * 1e3 / (spc * ke),
               "bytes_per_call": {"h2d": voxels, "d2h": voxels}, "numa": numa,
               "note": how + f", {spc} steps (one fused pass) per call: H2D of the whole grid, step kernels and D2H of "
                       "the result overlap chunk by chunk; wall clock" + (", max over ranks" if world_size > 1 else ""),
               "one_step_per_call": {"value": voxels * ke / dt1, "ms_per_step": dt1 * 1e3 / ke,
                                     "h2d_bytes_per_step": voxels, "d2h_bytes_per_step": voxels}}
        if spc == 4:
            e2e["two_steps_per_call"] = {"value": voxels * 2 * ke / dt2, "ms_per_step": dt2 * 1e3 / (2 * ke),
                                         "h2d_bytes_per_step": voxels // 2, "d2h_bytes_per_step": voxels // 2}
        if world_size == 1:
            # the same call for a grid the host keeps packed in the checkpoint encoding (2 bits per voxel): a quarter of
            # the bytes cross PCIe.  Reported beside the uint8 figure, not instead of it.
            pk = torch.empty((voxels // 4,), dtype=torch.uint8, pin_memory=True)
            pv = pk.numpy()
            while w.step_index % spc:
                w.step_host(hv, hv, 1)
            w.download_packed(pv)                        # the device holds what the last call returned
            w.step_host_packed(pv, pv, spc)
            t0 = time.perf_counter()
            for _ in range(ke):
                w.step_host_packed(pv, pv, spc)
            dtp = time.perf_counter() - t0
            e2e["packed_host_grid"] = {"value": voxels * spc * ke / dtp, "ms_per_step": dtp * 1e3 / (spc * ke),
                                       "h2d_bytes_per_step": voxels // (4 * spc), "d2h_bytes_per_step": voxels // (4 * spc),
                                       "steps_per_call": spc,
                                       "note": "fs3d_step_host_packed: the host holds 2 bits per voxel (checkpoint encoding), chunks are "
                                               f"unpacked / packed on the device; {spc} steps per call"}
            del pk
        del host
        stepper.close()

    extra = {}
    if args.big_steps > 0 and n != args.big_size:
        extra[f"size_{args.big_size}"] = big_grid_block(args, world_size, rank, hbm_peak)
    if args.m8_steps > 0 and world_size == 1:
        extra["materials8"] = materials8_block(args, hbm_peak)

    if rank != 0:
        if world_size > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    if world_size == 1 and not args.no_cpu:
        threads, omp_env, affinity = host_threads()
        nz_sample = sample_planes(n, threads, 1 << 28)            # >= 256 Mi voxels per step: 10-30 s in all
        rate, cores, dt, sample = cpu_oracle_rate(n, n, nz_sample, args.cpu_steps)
        cpu = {"value": rate, "unit": "voxel-updates/s", "cores": cores, "kind": "port", "sample": sample,
               "seconds": dt, "omp_num_threads_env": omp_env, "affinity_cpus": affinity,
               "note": "builder-written CPU oracle of the same schedule (the reference has no CPU update, SURVEY.md §0), "
                       "OpenMP over (plane, block-row) work items"}

    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
            # the dominant kernel of the N = 1 run: the four-step kernel where fs3d_step uses it (rows of 1024 / 2048 / 4096 voxels)
            traffic = (tj.get(f"{n}_fused4") if n in (1024, 2048, 4096) else tj.get(f"{n}_fused")) if world_size == 1 else None

    line = {
        "metric": METRIC, "value": value, "unit": "voxel-updates/s", "n_gpus": world_size, "steps": K, "warmup": Wm,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"{n}^3 MIXED_NOISE scene (BASELINE configs[3]: 2048^3 mixed sand/water/stone), "
                               f"generated on device, skipping off",
                   "grid": [n, n, n],
                   "parallelism": ("single GPU" if world_size == 1 else
                                   (f"z-slabs x{world_size}, four steps per pass, two ghost planes per side delivered over NVLink peer memory "
                                    f"by a kernel behind each pass" if getattr(sw, "fuse4", False) else
                                    f"z-slabs x{world_size}, halo pushed over NVLink peer memory inside the step kernel") if p2p else
                                   f"z-slabs x{world_size}, halo planes exchanged with NCCL send/recv on a side stream"),
                   "l2": "inputs larger than L2 (grid %.1f GiB per buffer vs 126 MB L2); no flush" % (voxels / 2**30),
                   "digest": hex(digest), "digest_step": total_steps, "digest_check": digest_check},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src,
                     # what the memory system really sustains: measured DRAM bytes per launch (ncu) over the measured launch time
                     "dram_gbs_from_traffic": (traffic * launches / (ms * 1e-3) / 1e9) if (traffic and world_size == 1) else None,
                     "hardware_frac": (traffic * launches / (ms * 1e-3) / 1e9 / hbm_peak) if (traffic and world_size == 1) else None,
                     "algorithmic_bytes_per_voxel_update": 2,
                     "kernel": ("fs3d::step4_kernel (four steps per launch)" if (world_size == 1 and n in (1024, 2048, 4096) and launches * 4 <= K + 3)
                                else "fs3d::step_kernel<NS=2> (two steps per launch)"),
                     "note": "achieved = 2 B x voxel-updates per launch / duration, per GPU. One launch advances every "
                             "voxel several steps while moving ~2 B per voxel, so frac can exceed 1: the real DRAM bytes "
                             "are `traffic` (hardware_frac = traffic / time / peak); the four-step kernel runs at ~0.8 of both of its "
                             "roofs (DRAM traffic and the integer ALU pipe, ncu: profiles/r02s_*); `single_step` is the unfused "
                             "kernel the 2 B/update roofline describes (0.96)"},
        "single_step": single,
        "two_steps_per_pass": two if world_size == 1 else None,
        "halo_wait": halo_wait if world_size > 1 else None,
        "extra": extra,
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-steps", type=int, default=24)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused-only", action="store_true", help="skip the unfused single-step measurement")
    ap.add_argument("--m8-steps", type=int, default=20, help="timed steps of the schedule-version-2 block (N = 1); 0 = skip it")
    ap.add_argument("--big-size", type=int, default=4096, help="grid edge of the extra large-grid block (BASELINE configs[4])")
    ap.add_argument("--big-steps", type=int, default=20, help="timed steps of the extra large-grid block; 0 = skip it")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
