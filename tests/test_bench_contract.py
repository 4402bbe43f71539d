"""bench.py contract (CPU side): the reference arm prints one JSON line with the agreed keys, the
synthetic scene generators are seeded, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ, NKSR_BENCH_WATCHDOG="280")
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=300, env=e)


def test_reference_arm_json_line():
    out = _run(["--impl", "reference", "--workload", "dev_outdoor_1M", "--steps", "1", "--warmup", "0",
                "--cpu-sample", "8000"])      # (cpu-sample < 50 K is honoured as given: test-sized)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "points/sec reconstruct()" and line["unit"] == "points/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["config"]["workload"] == "dev_outdoor_1M"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "crop" in cb["sample"] and cb["nnz"] > 0 and cb["pcg_iterations"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    out = _run(["--impl", "reference", "--workload", "dev_outdoor_1M", "--steps", "1", "--warmup", "0"],
               env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_scene_generators_are_seeded_and_shaped():
    sys.path.insert(0, ROOT)
    import bench
    a, sa = bench.make_cloud("dev_outdoor_1M", 4, 0, points=20000)
    b, sb = bench.make_cloud("dev_outdoor_1M", 4, 0, points=20000)
    c, _ = bench.make_cloud("dev_outdoor_1M", 5, 0, points=20000)
    assert a.shape == (20000, 3) and sa.shape == (20000, 3) and a.dtype.is_floating_point
    assert np.array_equal(a.numpy(), b.numpy()) and not np.array_equal(a.numpy(), c.numpy())
    t1, _ = bench.make_cloud("dev_outdoor_1M", 4, 1, points=20000)
    assert np.allclose(t1.numpy()[:, 0] - a.numpy()[:, 0], 200.0, atol=1e-3)      # tiles are shifted copies
    i, si = bench.make_cloud("cfg3_indoor_1M", 3, 0, points=5000)
    assert i.shape == (5000, 3) and float(i[:, 2].max()) < 3.1 and float(i[:, 2].min()) > -0.1
