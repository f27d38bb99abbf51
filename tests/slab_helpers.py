"""Test-only helpers: an oracle-backed engine for SlabWorld (so the partition / halo-exchange /
reduction host logic runs on CPU under gloo) and a spawn wrapper."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STONE = 3


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class OracleSlabEngine:
    """Same protocol as fallingsand3d_b200.slab.CudaSlabEngine, computed by the CPU oracle."""

    def __init__(self, nx, ny, nz, seed, z_begin, z_end):
        from oracle import oracle
        self.o = oracle
        self.nx, self.ny, self.nz, self.seed = nx, ny, nz, seed
        self.zb, self.ze = z_begin, z_end
        nzl = z_end - z_begin
        self.buf = [np.zeros((nzl + 2, ny, nx), np.uint8) for _ in range(2)]
        for b in self.buf:
            b[0] = STONE
            b[-1] = STONE
        self.cur = 0
        self.t = 0
        self.ns = 1

    def pass_steps(self, ns):
        self.ns = ns

    def halo_tensors(self, back):
        b = self.buf[self.cur ^ 1 if back else self.cur]
        return (torch.from_numpy(b[1]), torch.from_numpy(b[-2]), torch.from_numpy(b[0]), torch.from_numpy(b[-1]))

    def before_exchange(self, after_edges):
        return _NullCtx()

    def after_exchange(self):
        pass

    def step_edges(self):
        src, dst = self.buf[self.cur], self.buf[self.cur ^ 1]
        ghosts = (dst[0].copy(), dst[-1].copy())
        dst[...] = src
        # a fused pass evolves the ghost planes locally for its second step: the (ghost, owned) ZY pair
        # is closed under steps 2k and 2k + 1, so no exchange is needed in between
        for k in range(self.ns):
            self.o.step_range(dst, self.nz, self.zb - 1, self.zb, self.ze, self.seed, self.t + k)
        # the step scribbles on ghost planes; they are refreshed by the exchange (or stay STONE)
        dst[0], dst[-1] = ghosts
        if self.zb == 0:
            dst[0] = STONE
        if self.ze == self.nz:
            dst[-1] = STONE

    def step_interior(self):
        pass

    def step_finish(self):
        self.cur ^= 1
        self.t += self.ns

    def sync(self):
        pass

    def generate(self, scene, seed):
        self.buf[self.cur][1:-1] = self.o.generate(self.nx, self.ny, self.nz, scene, seed, self.zb, self.ze)

    def upload(self, a):
        self.buf[self.cur][1:-1] = a

    def download(self):
        return self.buf[self.cur][1:-1].copy()

    def digest(self):
        return self.o.digest(np.ascontiguousarray(self.buf[self.cur][1:-1]), self.zb)

    def histogram(self):
        return self.o.histogram(self.buf[self.cur][1:-1])

    def save(self, path):
        from fallingsand3d_b200 import checkpoint
        checkpoint.write(path, self.buf[self.cur][1:-1], nz=self.nz, z_begin=self.zb, step=self.t, seed=self.seed)

    def load(self, path):
        from fallingsand3d_b200 import checkpoint
        h, g = checkpoint.read(path)
        assert (h["nx"], h["ny"], h["nz"], h["z_begin"], h["z_end"]) == (self.nx, self.ny, self.nz, self.zb, self.ze)
        self.buf[self.cur][1:-1] = g
        self.t, self.seed = int(h["step"]), int(h["seed"])
        return self.t, self.seed

    def reduce_device(self):
        return torch.device("cpu")

    def close(self):
        pass


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(rank, world_size, port, fn, args, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        out = fn(rank, world_size, *args)
        q.put((rank, "ok", out))
    except Exception as e:  # pragma: no cover - reported to the parent
        import traceback
        q.put((rank, "err", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def run_ranks(world_size, fn, *args):
    """Runs fn(rank, world_size, *args) in world_size gloo processes; returns results by rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_entry, args=(r, world_size, port, fn, args, q)) for r in range(world_size)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, status, out = q.get(timeout=300)
        if status != "ok":
            for p in procs:
                p.terminate()
            raise AssertionError(f"rank {rank} failed:\n{out}")
        res[rank] = out
    for p in procs:
        p.join(timeout=60)
    return [res[r] for r in range(world_size)]
