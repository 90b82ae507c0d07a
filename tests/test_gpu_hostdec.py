"""GPU: the drop-in proof.  The reference's own vpxdec tool, linked against the reference's
host parser with the B200 seams (hostdec/_build/vpxdec_b200), decodes the golden IVF streams
through vpx_codec_decode / vpx_codec_get_frame and must print exactly the per-frame MD5s the
unmodified reference decoder printed (tests/golden/*.md5)."""
import os
import subprocess
import tempfile

import pytest

from conftest import CASES, GOLD, ROOT

pytestmark = pytest.mark.gpu

VPXDEC_B200 = os.path.join(ROOT, "hostdec", "_build", "vpxdec_b200")


@pytest.mark.parametrize("name", CASES)
def test_vpxdec_b200_md5_matches_reference(gpu_lib, name):
    assert os.path.exists(VPXDEC_B200), "hostdec is not built (run __graft_entry__.build() where the reference is present)"
    want = open(os.path.join(GOLD, name + ".md5")).read().split()
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ)
        env.pop("VP8B200_NO_DEVICE", None)
        out = subprocess.run([VPXDEC_B200, "--md5", "--i420", "-o", os.path.join(tmp, "f-%4.i420"),
                              os.path.join(GOLD, name + ".ivf")], env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    got = [l.split()[0] for l in out.stdout.splitlines() if l.strip()]
    assert got == want


def test_multi_stream_driver_runs(gpu_lib):
    """b200bench: 6 decoder instances on 3 threads over two golden clips."""
    exe = os.path.join(ROOT, "hostdec", "_build", "b200bench")
    assert os.path.exists(exe)
    out = subprocess.run([exe, "--threads", "3", "--streams", "6", "--repeat", "2", "--sum",
                          os.path.join(GOLD, "cif_p0.ivf"), os.path.join(GOLD, "qcif_arf.ivf")],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    import json
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["frames"] == 2 * 3 * (30 + 30) and r["checksum"] > 0 and r["kernel_launches"] > 0
    assert r["h2d_bytes"] > 0 and r["d2h_bytes"] > 0
