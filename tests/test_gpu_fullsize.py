"""GPU: BASELINE.json's full-size configurations, end to end through the drop-in decoder
(reference vpxdec + host parser + B200 seams): C2 1280x720 profile 0, C3 1920x1080 profile 3
(bilinear, full pixel), C3b 1920x1080 profile 1 (bilinear + SIMPLE loop filter), C4 3840x2160
with 8 token partitions + segmentation, and a sample of the C5 1080p bench clips.  Every
frame's MD5 must equal what the unmodified reference decoder printed (streams/<name>.md5,
written by tools/make_streams.py in the build container)."""
import glob
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

STREAMS = os.path.join(ROOT, "streams")
VPXDEC_B200 = os.path.join(ROOT, "hostdec", "_build", "vpxdec_b200")
NAMES = ["c2_720p", "c3_1080p_p3", "c3b_1080p_p1", "c4_2160p", "c5_1080p_s100", "c5_1080p_s131", "c5_1080p_s163"]


@pytest.mark.parametrize("name", NAMES)
def test_full_size_stream_md5(gpu_lib, name, tmp_path):
    ivf = os.path.join(STREAMS, name + ".ivf")
    if not os.path.exists(ivf):
        pytest.skip("streams/ not generated (tools/make_streams.py needs the reference encoder)")
    assert os.path.exists(VPXDEC_B200), "hostdec is not built"
    want = open(ivf[:-4] + ".md5").read().split()
    out = subprocess.run([VPXDEC_B200, "--md5", "--i420", "-o", str(tmp_path / "f-%4.i420"), ivf],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    got = [l.split()[0] for l in out.stdout.splitlines() if l.strip()]
    assert len(got) == len(want) and got == want, "%s: first mismatch at frame %d" % (
        name, next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), -1))
