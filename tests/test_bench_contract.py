"""CPU: the reference arm of bench.py (`--impl reference`) honours the driver's contract - one
JSON line from rank 0 with the metric / unit / config of the B200 arm, `impl`, `cpu_baseline`
and an `e2e` object; other ranks print nothing and exit 0.  The arm times the UNMODIFIED
reference decoder (oracle/_ref/refbench) on the host cores, so it runs without a GPU."""
import glob
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REFBENCH = os.path.join(ROOT, "oracle", "_ref", "refbench")
CLIPS = glob.glob(os.path.join(ROOT, "streams", "c5_1080p_s*.ivf"))
needs = pytest.mark.skipif(not (os.path.exists(REFBENCH) and CLIPS),
                           reason="oracle/_ref or streams/ not built (needs the reference sources)")


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--streams", "4",
                           "--steps", "1", "--warmup", "0"], env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


@needs
def test_reference_arm_prints_one_contract_line():
    out = _run({})
    assert out.returncode == 0, out.stderr[-400:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["config"]["workload"] == "c5_64x1080p"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


@needs
def test_reference_arm_other_ranks_are_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""
