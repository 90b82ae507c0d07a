"""CPU: the oracle restatement replays the committed record dumps and must reproduce the
per-frame MD5s the reference's own vpxdec (generic C) printed for the same IVF streams."""
import pytest

from conftest import CASES, load_case
from oracle_lib import OracleDecoder
from vp8b200 import frames


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_md5(name):
    rec, md5s = load_case(name)
    geo = frames.Geometry(rec.coded_width, rec.coded_height)
    dec = OracleDecoder(rec.coded_width, rec.coded_height, rec.n_fb)
    shown = 0
    for i, fr in enumerate(rec.frames):
        dec.frame(fr)
        if fr.show_frame:
            got = frames.md5_hex(geo.i420(dec.fb(fr.fb_show), rec.display_width, rec.display_height))
            assert got == md5s[shown], "%s frame %d" % (name, i)
            shown += 1
    assert shown == len(md5s)
