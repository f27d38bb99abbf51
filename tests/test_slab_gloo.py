"""CPU tests (gloo, world_size 2 and 3) of the N>1 host logic in fallingsand3d_b200/slab.py:
z-slab partition, one-plane halo exchange per step, global digest/histogram reductions.  The
compute engine is the CPU oracle here; the same SlabWorld drives libfs3d on the GPUs."""
import numpy as np
import pytest

from fallingsand3d_b200.slab import slab_bounds
from tests.slab_helpers import OracleSlabEngine, run_ranks


def test_slab_bounds_cover_and_prefer_even_boundaries():
    for nz in (2, 3, 7, 10, 64, 255, 2048):
        for ws in (1, 2, 3, 4, 8):
            if ws > nz:
                continue
            b = slab_bounds(nz, ws)
            assert b[0][0] == 0 and b[-1][1] == nz
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            assert all(lo < hi for lo, hi in b)
    assert slab_bounds(2048, 8) == [(256 * i, 256 * (i + 1)) for i in range(8)]
    assert all(lo % 2 == 0 for lo, _ in slab_bounds(250, 4))


def _worker(rank, world_size, dims, scene, seed, steps, chunk):
    from fallingsand3d_b200.slab import SlabWorld
    nx, ny, nz = dims
    sw = SlabWorld(nx, ny, nz, seed=seed, engine_factory=OracleSlabEngine)
    sw.generate(scene, 4)
    h0 = sw.histogram()
    digests = [sw.digest()]
    for _ in range(0, steps, chunk):
        sw.step(chunk)            # chunk >= 2 takes the fused 2-step passes (one exchange per pass)
        digests.append(sw.digest())
    assert np.array_equal(sw.histogram(), h0)
    whole = sw.gather()
    return digests, whole, [int(v) for v in h0[:4]], sw.exchanges, (sw.z_begin, sw.z_end)


@pytest.mark.parametrize("world_size,dims,chunk", [(2, (32, 12, 10), 1), (3, (64, 9, 11), 3), (2, (32, 8, 3), 1),
                                                   (2, (32, 12, 10), 4), (3, (64, 10, 13), 2)])
def test_slab_world_matches_whole_grid_oracle(oracle, world_size, dims, chunk):
    nx, ny, nz = dims
    steps, seed, scene = 12, 21, 3
    out = run_ranks(world_size, _worker, dims, scene, seed, steps, chunk)
    g = oracle.generate(nx, ny, nz, scene, 4)
    want = [oracle.digest(g)]
    for t in range(steps):
        oracle.step(g, seed, t)
        if (t + 1) % chunk == 0:
            want.append(oracle.digest(g))
    passes = {1: steps, 2: steps // 2, 3: (steps // 3) * 2, 4: steps // 2}[chunk]   # 3 = a pair + a single
    for rank, (digests, whole, h, exchanges, zr) in enumerate(out):
        assert digests == want, f"rank {rank}"
        assert np.array_equal(whole, g)
        assert h == [int(v) for v in oracle.histogram(oracle.generate(nx, ny, nz, scene, 4))[:4]]
        assert exchanges == passes + 2         # one per pass + the refresh at creation + the one after generate
    assert [o[4] for o in out] == slab_bounds(nz, world_size)


def _step_host_worker(rank, world_size, dims, seed):
    from fallingsand3d_b200.slab import SlabWorld
    from oracle import oracle
    nx, ny, nz = dims
    sw = SlabWorld(nx, ny, nz, seed=seed, engine_factory=OracleSlabEngine)
    full = oracle.generate(nx, ny, nz, 4, 2)
    host = full[sw.z_begin:sw.z_end].copy()      # a copy: the slice itself is contiguous and would alias `full`
    out = np.empty_like(host)
    t = 0
    for n in (2, 1, 1, 2):                       # host slab in, stepped host slab out, every call
        sw.step_host(host, out, n)
        oracle.run(full, seed, t, n)
        t += n
        assert np.array_equal(out, full[sw.z_begin:sw.z_end]), f"rank {rank} after step {t}"
        host, out = out, host
    sw.step(3)                                   # and ordinary stepping continues from the stepped state
    oracle.run(full, seed, t, 3)
    assert np.array_equal(sw.download(), full[sw.z_begin:sw.z_end])
    return sw.step_index


@pytest.mark.parametrize("world_size,dims", [(2, (32, 10, 9)), (3, (64, 8, 12))])
def test_step_host_on_host_resident_slabs(world_size, dims):
    # the transport-independent fallback of SlabWorld.step_host (upload + step + download per rank); the GPU
    # ranks stream their slabs instead (tests/run_slab_ranks.py)
    assert run_ranks(world_size, _step_host_worker, dims, 17) == [9] * world_size


def _ckpt_worker(rank, world_size, dims, seed, base):
    from fallingsand3d_b200.slab import SlabWorld
    from fallingsand3d_b200 import checkpoint
    nx, ny, nz = dims
    sw = SlabWorld(nx, ny, nz, seed=seed, engine_factory=OracleSlabEngine)
    sw.generate(4, 3)
    sw.step(5)                                   # odd step index: the resumed run must keep the schedule phase
    sw.save(base)
    h, g = checkpoint.read(sw.rank_path(base))
    assert (h["z_begin"], h["z_end"], h["nz"], h["step"], h["seed"]) == (sw.z_begin, sw.z_end, nz, 5, seed)
    assert np.array_equal(g, sw.download())
    sw.step(7)
    want = sw.digest()
    sw.step(2)
    sw.load(base)
    assert sw.step_index == 5
    sw.step(7)
    return sw.digest() == want, sw.rank_path(base)


def test_per_rank_checkpoints_resume_bit_identically(tmp_path):
    out = run_ranks(3, _ckpt_worker, (32, 8, 11), 9, str(tmp_path / "w"))
    assert all(ok for ok, _ in out)
    assert len({p for _, p in out}) == 3          # one file per rank, named by its plane range


def test_u64_allreduce_wraps_like_the_digest():
    out = run_ranks(2, _wrap_worker)
    assert out[0] == out[1] == [(2 ** 64 - 5 + 2 ** 63 + 11) % 2 ** 64, 7]


def _wrap_worker(rank, world_size):
    from fallingsand3d_b200.slab import SlabWorld
    sw = SlabWorld(32, 4, 4, engine_factory=OracleSlabEngine)
    vals = np.array([2 ** 64 - 5, 3] if rank == 0 else [2 ** 63 + 11, 4], dtype=np.uint64)
    return [int(v) for v in sw._allreduce_u64(vals)]
