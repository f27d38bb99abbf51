"""Launched under torchrun on >= 2 GPUs (see test_slab_nccl_gpu.py): SlabWorld over NCCL must
reproduce the single-GPU world's digest step by step, and the oracle's on a small grid."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fallingsand3d_b200 as fs3d  # noqa: E402
from fallingsand3d_b200.slab import SlabWorld  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    SK = fs3d.FLAG_SKIP_SETTLED
    cases = [(64, 24, 18, 3, 12, True, 0), (64, 24, 18, 3, 12, False, 0), (2048, 32, 24, 4, 10, True, 0),
             (2048, 32, 24, 4, 10, False, 0), (256, 256, 256, 2, 40, True, 0), (4096, 16, 20, 3, 6, True, 0),
             (1024, 48, 40, 4, 16, True, 0), (2048, 64, 32, 4, 12, True, 0),      # four-step passes across ranks (slab.fuse4)
             (4096, 24, 32, 4, 12, True, 0),
             (64, 96, 48, 1, 60, True, SK), (64, 96, 48, 1, 30, False, SK), (4096, 70, 16, 4, 12, True, SK)]
    for (nx, ny, nz, scene, steps, p2p, flags) in cases:
        sw = SlabWorld(nx, ny, nz, seed=5, p2p=p2p, flags=flags)
        assert sw.p2p == p2p
        nzl = nz // dist.get_world_size()
        if p2p and nx in (1024, 2048, 4096) and not flags and nz % dist.get_world_size() == 0 and nzl >= 4 and nzl % 2 == 0:
            assert sw.fuse4, "every rank's slab supports four-step passes here"
        sw.generate(scene, 3)
        ref = None
        if rank == 0:
            ref = fs3d.VoxelWorld(nx, ny, nz, seed=5)
            ref.generate(scene, 3)
        t = 0
        for chunk in [1, 2, 3] * steps:
            if t >= steps:
                break
            sw.step(chunk)
            t += chunk
            d = sw.digest()
            if rank == 0:
                ref.step(chunk)
                if ref.digest() != d:
                    ok = False
                    print(f"MISMATCH {nx}x{ny}x{nz} p2p={p2p} step {t}", flush=True)
                    break
        if ok and p2p:
            # many passes back to back with no host synchronisation in between (four-step passes where the slabs allow)
            sw.step(50)
            d = sw.digest()
            if rank == 0:
                ref.step(50)
                if ref.digest() != d:
                    ok = False
                    print(f"MISMATCH {nx}x{ny}x{nz} p2p after 50 async steps", flush=True)
        # every rank marches its slab; rank 0 composites (p2p: in the kernels over peer memory, else NCCL gather)
        cam = dict(pos=(0.2, -0.3, -1.3), yaw_deg=20.0, aspect=16.0 / 9.0, width=192, height=108, mode=fs3d.RM_VOXELS)
        for _ in range(2):
            img = sw.raymarch(**cam)
            if rank == 0 and not np.array_equal(img, ref.raymarch(**cam)):
                ok = False
                print(f"MISMATCH {nx}x{ny}x{nz} p2p={p2p} raymarch composite", flush=True)
        # per-rank checkpoints: save, run on, load, run on again -> same digest as the uninterrupted run
        import tempfile
        base = [tempfile.mkdtemp(prefix="fs3d_ckpt_") if rank == 0 else None]
        dist.broadcast_object_list(base, src=0)
        ck = os.path.join(base[0], "w")
        sw.save(ck)
        t_saved = sw.step_index
        sw.step(5)
        d_after = sw.digest()
        sw.step(3)
        sw.load(ck)
        if sw.step_index != t_saved:
            ok = False
        sw.step(5)
        if sw.digest() != d_after:
            ok = False
            print(f"MISMATCH {nx}x{ny}x{nz} p2p={p2p} after checkpoint resume", flush=True)
        if rank == 0:
            ref.step(5)
            ok = ok and ref.digest() == d_after
        h = sw.histogram()
        if rank == 0:
            ok = ok and np.array_equal(h, ref.histogram())
            ref.close()
        sw.close()
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        if not ok:
            break
    # end-to-end step of host-resident slabs: every rank streams its planes (fs3d_slab_step_host); checked against a
    # whole-grid world that each rank runs on its own GPU
    for (nx, ny, nz, p2p) in [(64, 40, 30, True), (2048, 24, 20, True), (64, 40, 30, False)]:
        sw = SlabWorld(nx, ny, nz, seed=11, p2p=p2p)
        with fs3d.VoxelWorld(nx, ny, nz, seed=11) as ref:
            ref.generate(fs3d.SCENE_MIXED_NOISE, 6)
            zb, ze = sw.z_begin, sw.z_end
            host = np.ascontiguousarray(ref.download()[zb:ze])
            out = np.empty_like(host)
            t = 0
            for n in (2, 1, 1, 2, 2, 1):
                if n == 2 and t % 2:
                    n = 1
                sw.step_host(host, out, n)
                ref.step(n)
                t += n
                if not np.array_equal(out, ref.download()[zb:ze]):
                    ok = False
                    print(f"MISMATCH step_host {nx}x{ny}x{nz} p2p={p2p} rank {rank} after step {t}", flush=True)
                    break
                host, out = out, host
            if p2p:
                try:                        # the ghost planes are stale now: the library refuses to step on them
                    sw.engine.world.step(1)
                    ok = False
                    print("fs3d_step accepted stale ghost planes after fs3d_slab_step_host", flush=True)
                except fs3d.Fs3dError as e:
                    ok = ok and e.code == -7
            sw.step(5)                      # ordinary stepping continues (halos are refreshed first)
            ref.step(5)
            if not np.array_equal(sw.download(), ref.download()[zb:ze]):
                ok = False
                print(f"MISMATCH after step_host + step {nx}x{ny}x{nz} p2p={p2p} rank {rank}", flush=True)
        sw.close()
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        if not ok:
            break
    # step -> upload -> step with one rank's GPU deliberately late: a pass only waits for the neighbours'
    # previous pass, so refresh_halos must drain and barrier BEFORE it stores into a neighbour's ghost plane
    if ok:
        nx, ny, nz = 256, 64, 24
        sw = SlabWorld(nx, ny, nz, seed=13, p2p=True)
        with fs3d.VoxelWorld(nx, ny, nz, seed=13) as ref:
            sw.generate(fs3d.SCENE_MIXED_NOISE, 4)
            ref.generate(fs3d.SCENE_MIXED_NOISE, 4)
            sw.step(2)
            if rank == dist.get_world_size() - 1:
                with torch.cuda.stream(sw.engine.stream):
                    torch.cuda._sleep(1_500_000_000)          # ~0.75 s of GPU time in front of this rank's next pass
            sw.step(2)
            ref.step(4)
            other = np.ascontiguousarray(ref.download()[::-1, :, ::-1])       # some other valid grid
            sw.upload(np.ascontiguousarray(other[sw.z_begin:sw.z_end]))
            ref.upload(other)
            sw.step(6)
            ref.step(6)
            if not np.array_equal(sw.download(), ref.download()[sw.z_begin:sw.z_end]):
                ok = False
                print(f"MISMATCH after step / upload / step with a delayed rank (rank {rank})", flush=True)
        sw.close()
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    if rank == 0:
        print("SLAB_NCCL_OK" if ok else "SLAB_NCCL_FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
