"""CPU (build container only): tools/refconfig.py - which writes the configuration headers
oracle/refbuild and hostdec compile the reference with - against the reference's OWN
`./configure --target=generic-gnu --disable-multithread` + `make vpx_rtcd.h`, run on a scratch
copy of /root/reference (that configuration is the generic-C ground truth).
Every CONFIG_/HAVE_/ARCH_ switch and every RTCD binding of the reference's build must be the
same in ours; the only tolerated differences are listed with their reason."""
import os
import re
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("VP8B200_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")

# CONFIG_PIC: we build shared objects (-fPIC); the switch only selects assembler variants,
#             and the generic target has no assembler.
ALLOWED_CONFIG_DIFF = {"CONFIG_PIC"}


def _defines(path, pattern):
    out = {}
    for line in open(path):
        m = re.match(pattern, line)
        if m:
            out[m.group(1)] = m.group(2)
    return out


def test_refconfig_equals_the_reference_configure(tmp_path):
    src = tmp_path / "ref"
    shutil.copytree(REF, src)
    subprocess.check_call(["chmod", "-R", "u+w", str(src)])
    subprocess.check_call(["./configure", "--target=generic-gnu", "--disable-multithread"], cwd=src,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["make", "vpx_rtcd.h"], cwd=src, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ours = tmp_path / "ours"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "refconfig.py"), REF, str(ours)],
                          stdout=subprocess.DEVNULL)
    pat = r"#define ((?:CONFIG|HAVE|ARCH)_\w+) (\w+)"
    a, b = _defines(src / "vpx_config.h", pat), _defines(ours / "vpx_config.h", pat)
    assert set(a) == set(b), (sorted(set(a) - set(b)), sorted(set(b) - set(a)))
    diff = {k for k in a if a[k] != b[k]}
    assert diff <= ALLOWED_CONFIG_DIFF, {k: (a[k], b[k]) for k in diff}
    # RTCD: every name the reference's build binds must be bound to the same implementation
    pat = r"#define (vp8_\w+) (vp8_\w+)"
    a, b = _defines(src / "vpx_rtcd.h", pat), _defines(ours / "vpx_rtcd.h", pat)
    assert len(a) > 80
    wrong = {k: (v, b.get(k)) for k, v in a.items() if b.get(k) != v}
    assert not wrong, wrong
    assert all(v.endswith("_c") for v in a.values())             # the generic C path, nothing else
    # names only ours defines must still be plain C bindings (encoder helpers outside the hot path)
    assert all(b[k] == k + "_c" for k in set(b) - set(a)), sorted(set(b) - set(a))
