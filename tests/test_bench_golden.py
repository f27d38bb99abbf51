"""CPU-side checks of the committed oracle digests of bench.py's workload (tests/golden/bench_digests.json)."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_bench_digest_file_is_consistent(oracle):
    with open(os.path.join(GOLDEN, "bench_digests.json")) as f:
        gold = json.load(f)
    assert gold["schedule_version"] == oracle.lib().fs3d_oracle_schedule_version()
    assert gold["dims"] == [2048, 2048, 2048] and gold["scene"] == 4
    # the step counts bench.py ends on with the driver's flags (--steps 20 --warmup 5 -> 6 + 20) and its defaults (4 + 100)
    assert "26" in gold["digests"] and "104" in gold["digests"]
    assert sum(gold["histogram"]) == 2048 ** 3
    # a thin slab of the same scene, regenerated here, carries the same cells the digest file was made from:
    # its digest of step 0 over planes [1020, 1028) must be reproducible (pins the scene generator, not the 8 GiB run)
    g = oracle.generate(2048, 2048, 2048, 4, 1, 1020, 1028)
    assert oracle.digest(g, 1020) == int(gold["slab_1020_1028_digest0"], 16)


def test_bench_golden_check_raises_on_mismatch():
    import bench
    with open(os.path.join(GOLDEN, "bench_digests.json")) as f:
        gold = json.load(f)
    ok = bench.golden_digest_check(2048, 26, int(gold["digests"]["26"], 16))
    assert "equals the CPU oracle" in ok
    assert "no oracle digest" in bench.golden_digest_check(2048, 27, 0)
    assert "no oracle digest" in bench.golden_digest_check(1024, 26, 0)
    try:
        bench.golden_digest_check(2048, 26, 1)
    except AssertionError:
        return
    raise AssertionError("a wrong digest must fail the bench")
