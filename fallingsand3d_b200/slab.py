"""SlabWorld — one process per GPU: z-slab decomposition with a one-plane halo exchange per step
over torch.distributed (NCCL over NVLink on the GPUs; gloo in the CPU tests of this host logic).

No reference counterpart: the reference is single-GPU, single-threaded and has no distributed
backend (SURVEY.md §5 "Distributed communication backend: None").  SURVEY.md §8(e) is the spec:
each rank owns a contiguous z-slab plus one ghost plane each side; per step the two edge planes
are computed first, sent to the z-neighbours while the interior is computed, and both sides of a
boundary-crossing ZY block are evaluated redundantly from identical data (global coordinates and
the counter hash make them agree), so no move message is needed.

The compute engine is pluggable so that the partition / exchange / reduction logic can be tested
on CPU: CudaSlabEngine drives libfs3d through the C ABI; tests inject an oracle-backed engine.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(nz, world_size):
    """Contiguous z-ranges, boundaries on even planes where possible (the same rule as fs3d_create)."""
    if world_size > nz:
        raise ValueError("more ranks than z-planes")
    bounds = []
    zb = 0
    for i in range(world_size):
        ze = nz * (i + 1) // world_size
        if i + 1 < world_size and (ze & 1):
            ze += 1                                   # prefer even boundaries
        ze = min(ze, nz - (world_size - 1 - i))       # leave a plane for every later rank
        ze = max(ze, zb + 1)
        if i + 1 == world_size:
            ze = nz
        bounds.append((zb, ze))
        zb = ze
    return bounds


class _DevPtr:
    """Zero-copy torch view of library-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


class CudaSlabEngine:
    """libfs3d slab world on the current CUDA device (fs3d_create_slab + fs3d_slab_* protocol)."""

    def __init__(self, nx, ny, nz, seed, z_begin, z_end, device, flags=0):
        from .world import VoxelWorld
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.world = VoxelWorld(nx, ny, nz, seed=seed, flags=flags, slab=(z_begin, z_end))
        self._views = {}
        h = self.world.slab_halo(0)
        self.stream = torch.cuda.ExternalStream(h.stream, device=self.device)
        self.comm_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.edges_done = torch.cuda.Event()

    def halo_tensors(self, back):
        h = self.world.slab_halo(back)
        key = (h.send_lo, h.send_hi, h.recv_lo, h.recv_hi)
        if key not in self._views:
            n = h.plane_bytes
            self._views[key] = tuple(torch.as_tensor(_DevPtr(p, n), device=self.device)
                                     for p in (h.send_lo, h.send_hi, h.recv_lo, h.recv_hi))
        return self._views[key]

    # stream hooks: the exchange runs on comm_stream, ordered after the edge kernels, and the next
    # step's kernels are ordered after it
    def before_exchange(self, after_edges):
        if after_edges:
            self.comm_stream.wait_event(self.edges_done)   # not the interior kernel enqueued behind it
        else:
            self.comm_stream.wait_stream(self.stream)
        return torch.cuda.stream(self.comm_stream)

    def after_exchange(self):
        self.stream.wait_stream(self.comm_stream)

    def pass_steps(self, ns):
        self.world.slab_pass_steps(ns)

    def ipc_export(self):
        return self.world.slab_ipc_export()

    def ipc_attach(self, lo, hi):
        self.world.slab_ipc_attach(lo, hi)

    def push_halos(self):
        self.world.slab_push_halos()

    def step_edges(self):
        self.world.slab_step_edges()
        self.edges_done.record(self.stream)

    def step_interior(self):
        self.world.slab_step_interior()

    def step_finish(self):
        self.world.slab_step_finish()

    def sync(self):
        self.world.sync()
        self.comm_stream.synchronize()

    def generate(self, scene, seed):
        self.world.generate(scene, seed)

    def upload(self, a):
        self.world.upload(a)

    def download(self):
        return self.world.download()

    def digest(self):
        return self.world.digest()

    def histogram(self):
        return self.world.histogram()

    def save(self, path):
        self.world.save(path)

    def load(self, path):
        """-> (step index, seed) the checkpoint restored."""
        self.world.load(path)
        return self.world.step_index, self.world.seed

    def reduce_device(self):
        return self.device

    def close(self):
        self.world.close()


class SlabWorld:
    """The rank-local piece of a global nx×ny×nz world, stepped in lock-step with the other ranks."""

    def __init__(self, nx, ny, nz, seed=1, flags=0, engine_factory=None, group=None, fuse=True, p2p=True):
        self.group = group
        self.fuse = fuse          # steps 2k, 2k+1 share one pass and ONE halo exchange
        self.p2p = False          # set below: fused halo push over peer memory instead of NCCL send/recv
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nx, self.ny, self.nz, self.seed = nx, ny, nz, seed
        self.bounds = slab_bounds(nz, self.world_size)
        self.z_begin, self.z_end = self.bounds[self.rank]
        if engine_factory is None:
            dev = torch.cuda.current_device()
            self.engine = CudaSlabEngine(nx, ny, nz, seed, self.z_begin, self.z_end, dev, flags)
        else:
            self.engine = engine_factory(nx, ny, nz, seed, self.z_begin, self.z_end)
        self.step_index = 0
        self.exchanges = 0
        if p2p and fuse and self.world_size > 1 and hasattr(self.engine, "ipc_export"):
            # plumbing only: swap CUDA IPC handles with the z-neighbours; afterwards the step kernel
            # itself moves the halos over NVLink (fs3d.h "fused halo push"), no collective per step
            blobs = [None] * self.world_size
            dist.all_gather_object(blobs, self.engine.ipc_export(), group=self.group)
            lo = blobs[self.rank - 1] if self.rank > 0 else None
            hi = blobs[self.rank + 1] if self.rank + 1 < self.world_size else None
            self.engine.ipc_attach(lo, hi)
            # four-step passes only if EVERY rank's slab supports them: the ranks must fuse alike
            can = torch.tensor([1 if self.engine.world.slab_can_fuse4() else 0], device=self.engine.reduce_device())
            dist.all_reduce(can, op=dist.ReduceOp.MIN, group=self.group)
            self.fuse4 = bool(can.item())
            self.engine.world.slab_allow_fuse4(self.fuse4)
            dist.barrier(group=self.group)
            self.p2p = True
        if self.world_size > 1:
            self.refresh_halos()      # ghost planes between ranks start as the neighbour's (EMPTY) plane, not STONE

    # ---- halo exchange of one buffer (back = the buffer the current step is writing) ----
    def _exchange(self, back):
        if self.world_size == 1:
            return
        send_lo, send_hi, recv_lo, recv_hi = self.engine.halo_tensors(back)
        ops = []
        lo, hi = self.rank - 1, self.rank + 1
        # post receives first; the order of sends/recvs is the same on every rank pair
        if lo >= 0:
            ops.append(dist.P2POp(dist.irecv, recv_lo, lo, self.group))
            ops.append(dist.P2POp(dist.isend, send_lo, lo, self.group))
        if hi < self.world_size:
            ops.append(dist.P2POp(dist.isend, send_hi, hi, self.group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, hi, self.group))
        with self.engine.before_exchange(after_edges=bool(back)):
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.engine.after_exchange()
        self.exchanges += 1

    def refresh_halos(self):
        """After generate/upload/set_cell: make the front buffer's ghost planes current."""
        if self.p2p:
            # A rank's pass k only waits for its neighbours' pass k - 1, so a neighbour may still be running its last
            # pass — reading the very ghost plane an upload / load (which flip buffers) is about to overwrite.  Every
            # rank therefore drains its stream and meets the others BEFORE anything is stored into a neighbour.
            self.engine.sync()
            dist.barrier(group=self.group)
            self.engine.push_halos()
            dist.barrier(group=self.group)
            self.exchanges += 1
            return
        self._exchange(back=0)

    def generate(self, scene, seed=None):
        self.engine.generate(scene, self.seed if seed is None else seed)
        self.refresh_halos()

    def upload(self, local_planes):
        self.engine.upload(local_planes)
        self.refresh_halos()

    def download(self):
        return self.engine.download()

    # ---- checkpoint: one file per rank (format in include/fs3d.h) ----
    def rank_path(self, path):
        return f"{path}.z{self.z_begin}-{self.z_end}"

    def save(self, path):
        """Every rank writes its slab to rank_path(path)."""
        self.engine.save(self.rank_path(path))
        if self.world_size > 1:
            dist.barrier(group=self.group)

    def load(self, path):
        """Every rank restores its slab, step index and seed from rank_path(path); halos are refreshed."""
        self.step_index, self.seed = self.engine.load(self.rank_path(path))
        if self.world_size > 1:
            self.refresh_halos()

    def step_host(self, host_in, host_out=None, n=1):
        """End-to-end step of a host-resident slab (pinned memory recommended): every rank streams its planes
        through its GPU with overlapped copies; only the two edge planes are exchanged first (peer memory)."""
        if not self.p2p:
            self.upload(host_in)
            self.step(n)
            out = host_in if host_out is None else host_out
            out[...] = self.download()
            return out
        w = self.engine.world
        w.slab_step_host_begin(host_in)
        dist.barrier(group=self.group)          # every ghost plane holds its neighbour's edge plane
        out = w.slab_step_host(host_in, host_out, n)
        # No second barrier: the next call's _begin stores into the ghost planes of the buffer this call WROTE
        # (the library flipped buffers), which no rank reads before it has passed the next call's barrier above —
        # and a rank only gets there after its own slab_step_host of this call has returned.
        self.step_index += int(n)
        self._halos_stale = True
        return out

    def step(self, n=1):
        if getattr(self, "_halos_stale", False):
            self._halos_stale = False
            self.refresh_halos()
        if self.p2p:
            self.engine.world.step(int(n))      # the library loops; halos move inside the kernels
            self.step_index += int(n)
            return
        left = int(n)
        while left > 0:
            ns = 2 if (self.fuse and left >= 2 and self.step_index % 2 == 0) else 1
            self.engine.pass_steps(ns)
            self.engine.step_edges()        # the two edge planes of the back buffer are final after this
            self.engine.step_interior()     # enqueue first so it overlaps with the exchange below
            self._exchange(back=1)
            self.engine.step_finish()
            self.step_index += ns
            left -= ns

    def sync(self):
        self.engine.sync()

    # ---- global reductions ----
    def _allreduce_u64(self, arr):
        if self.world_size == 1:
            return arr
        # split into 32-bit halves so that a float-free integer SUM cannot overflow int64
        a = np.asarray(arr, dtype=np.uint64)
        parts = np.stack([(a & np.uint64(0xFFFFFFFF)).astype(np.int64), (a >> np.uint64(32)).astype(np.int64)])
        t = torch.from_numpy(parts).to(self.engine.reduce_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        p = t.cpu().numpy().astype(np.uint64)
        return (p[0] + (p[1] << np.uint64(32))).astype(np.uint64)   # wraps mod 2^64, like the digest sum

    def digest(self):
        return int(self._allreduce_u64(np.array([self.engine.digest()], dtype=np.uint64))[0])

    def histogram(self):
        return self._allreduce_u64(self.engine.histogram())

    def gather(self):
        """Whole grid on every rank (tests / small worlds only)."""
        local = torch.from_numpy(self.download())
        if self.world_size == 1:
            return local.numpy()
        outs = [None] * self.world_size
        dist.all_gather_object(outs, local.numpy(), group=self.group)
        return np.concatenate(outs, axis=0)

    def raymarch(self, **cam):
        """Every rank marches its own slab (fs3d_raymarch_depth); rank 0 keeps, per pixel, the colour of
        the nearest hit.  Returns the (H, W, 4) uint8 image on rank 0, else None."""
        if self.world_size == 1:
            return self.engine.world.raymarch(**cam)
        if self.p2p:
            return self._raymarch_fused(**cam)
        img, depth = self.engine.world.raymarch(with_depth=True, **cam)
        dev = self.engine.reduce_device()
        timg = torch.from_numpy(img).to(dev)
        tdep = torch.from_numpy(depth).to(dev)
        imgs = [torch.empty_like(timg) for _ in range(self.world_size)] if self.rank == 0 else None
        deps = [torch.empty_like(tdep) for _ in range(self.world_size)] if self.rank == 0 else None
        dist.gather(timg, imgs, dst=0, group=self.group)
        dist.gather(tdep, deps, dst=0, group=self.group)
        if self.rank != 0:
            return None
        d = torch.stack(deps)                       # (ranks, H, W)
        best = torch.argmin(d, dim=0)               # first minimum: ties cannot differ in colour
        allimg = torch.stack(imgs)                  # (ranks, H, W, 4)
        out = torch.gather(allimg, 0, best[None, :, :, None].expand(1, *allimg.shape[1:]))[0]
        return out.cpu().numpy()

    def _raymarch_fused(self, width=850, height=450, **cam):
        """March + composite over peer memory: every rank's kernel stores (t, rgba) into its slot of rank
        0's frame over NVLink; rank 0 takes the per-pixel minimum.  No host copy, no collective."""
        w = self.engine.world
        if getattr(self, "_frame_dims", None) != (width, height):
            blob = [w.frame_export(width, height, self.world_size) if self.rank == 0 else None]
            dist.broadcast_object_list(blob, src=0, group=self.group)
            w.frame_attach(blob[0], self.rank)
            self._frame_dims = (width, height)
        import time
        t0 = time.perf_counter()
        dist.barrier(group=self.group)          # rank 0 has resolved the previous frame: slots may be overwritten
        t1 = time.perf_counter()
        w.raymarch_to_frame(**cam)
        w.sync()
        t2 = time.perf_counter()
        dist.barrier(group=self.group)          # every slot is complete
        t3 = time.perf_counter()
        img = w.frame_resolve(width, height) if self.rank == 0 else None
        t4 = time.perf_counter()
        # where a frame's wall time goes on this rank (ms): tools/run_configs.py config 5 reports rank 0's
        self.last_frame_ms = {"barrier_before": (t1 - t0) * 1e3, "march_kernel_and_sync": (t2 - t1) * 1e3,
                              "barrier_after": (t3 - t2) * 1e3, "resolve_and_d2h": (t4 - t3) * 1e3}
        return img

    def close(self):
        self.engine.close()
