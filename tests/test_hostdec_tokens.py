"""CPU: the fused token reader of the host path (hostdec/vp8b200_tokens.c).

Two independent checks:
  * round trip - a bool ENCODER written here (RFC 6386 section 7.3 / 13) turns random
    coefficient blocks into a token partition; the C reader must give back the same
    coefficients, block mask, eobtotal and entropy contexts, macroblock after macroblock from
    one continuous partition (so the carried decoder state is covered too);
  * against the reference - when the patched host decoder is built (hostdec/_build, needs the
    reference sources, i.e. the build container), whole record dumps produced with the
    reference's vp8_decode_mb_tokens (VP8B200_TOKENS=ref) and with the fused reader must be
    byte-identical, and equal to the committed golden .rec fixtures.
"""
import ctypes
import lzma
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import CASES, GOLD, ROOT

SRC = os.path.join(ROOT, "hostdec", "vp8b200_tokens.c")
VPXDEC_B200 = os.path.join(ROOT, "hostdec", "_build", "vpxdec_b200")

ZIGZAG = [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]
BAND = [0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7]
CATS = [(5, [159]), (7, [165, 145]), (11, [173, 148, 140]), (19, [176, 155, 140, 135]),
        (35, [180, 157, 141, 134, 130]), (67, [254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129])]


class BoolEncoder:
    """RFC 6386 section 7.3."""

    def __init__(self):
        self.out = bytearray()
        self.range, self.bottom, self.bit_count = 255, 0, 24

    def _carry(self):
        i = len(self.out) - 1
        while i >= 0 and self.out[i] == 255:
            self.out[i] = 0
            i -= 1
        assert i >= 0
        self.out[i] += 1

    def put(self, bit, prob):
        split = 1 + (((self.range - 1) * int(prob)) >> 8)
        if bit:
            self.bottom += split
            self.range -= split
        else:
            self.range = split
        while self.range < 128:
            self.range <<= 1
            if self.bottom & (1 << 31):
                self._carry()
            self.bottom = (self.bottom << 1) & 0xFFFFFFFF
            self.bit_count -= 1
            if self.bit_count == 0:
                self.out.append(self.bottom >> 24)
                self.bottom &= (1 << 24) - 1
                self.bit_count = 8

    def finish(self):
        c, v = self.bit_count, self.bottom
        if v & (1 << (32 - c)):
            self._carry()
        v = (v << (c & 7)) & 0xFFFFFFFF
        c >>= 3
        for _ in range(c):
            v = (v << 8) & 0xFFFFFFFF
        for _ in range(4):
            self.out.append(v >> 24)
            v = (v << 8) & 0xFFFFFFFF
        return bytes(self.out)


def encode_block(enc, probs_t, ctx, first, zz):
    """zz: 16 coefficients in zigzag order.  Returns the end-of-block position as the
    reference reports it (detokenize.c:347-349: 15 when position 15 is coded)."""
    last = max([i for i in range(first, 16) if zz[i]] + [first - 1])
    c, after_zero = first, False
    while c < 16:
        p = probs_t[BAND[c]][ctx]
        if not after_zero:
            if c > last:
                enc.put(0, p[0])
                return c
            enc.put(1, p[0])
        v = int(zz[c])
        if v == 0:
            enc.put(0, p[1])
            ctx, after_zero = 0, True
            c += 1
            continue
        enc.put(1, p[1])
        a = abs(v)
        if a == 1:
            enc.put(0, p[2])
        else:
            enc.put(1, p[2])
            if a <= 4:
                enc.put(0, p[3])
                if a == 2:
                    enc.put(0, p[4])
                else:
                    enc.put(1, p[4])
                    enc.put(a - 3, p[5])
            else:
                enc.put(1, p[3])
                cat = max(i for i, (base, _) in enumerate(CATS) if a >= base)
                base, xp = CATS[cat]
                if cat < 2:
                    enc.put(0, p[6])
                    enc.put(cat, p[7])
                else:
                    enc.put(1, p[6])
                    enc.put((cat - 2) >> 1, p[8])
                    enc.put((cat - 2) & 1, p[9 + ((cat - 2) >> 1)])
                extra = a - base
                for k, pr in enumerate(xp):
                    enc.put((extra >> (len(xp) - 1 - k)) & 1, pr)
        enc.put(1 if v < 0 else 0, 128)
        ctx, after_zero = (1 if a == 1 else 2), False
        if c == 15:
            return 15
        c += 1
    return 15


class BoolDec(ctypes.Structure):
    _fields_ = [("buf", ctypes.c_void_p), ("buf_end", ctypes.c_void_p), ("value", ctypes.c_uint64),
                ("count", ctypes.c_int), ("range", ctypes.c_uint)]


@pytest.fixture(scope="module")
def tokens_lib():
    tmp = tempfile.mkdtemp(prefix="vp8b200_tok_")
    so = os.path.join(tmp, "libtok.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-Wall", "-Werror", "-I" + os.path.dirname(SRC), SRC, "-o", so],
                   check=True)
    lib = ctypes.CDLL(so)
    lib.vp8b200_decode_mb_tokens.restype = ctypes.c_int
    lib.vp8b200_decode_mb_tokens.argtypes = [ctypes.POINTER(BoolDec), ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.POINTER(ctypes.c_uint32)]
    return lib


def random_block(rng, first, density):
    zz = np.zeros(16, np.int64)
    if rng.random() < density:
        n = int(rng.integers(first + 1, 17))
        mags = rng.choice([1, 1, 1, 2, 3, 4, 5, 6, 7, 10, 11, 18, 19, 34, 35, 66, 67, 500, 2048 + 66], size=16)
        zz[first:n] = mags[first:n] * rng.choice([-1, 1], size=16)[first:n] * (rng.random(16)[first:n] < 0.6)
    return zz


@pytest.mark.parametrize("seed,density,pad", [(1, 0.15, 16), (2, 0.6, 16), (3, 1.0, 16), (4, 0.0, 16),
                                              (5, 0.6, 0), (6, 0.02, 0)])   # pad 0: byte-wise tail of the refill
def test_token_reader_round_trip(tokens_lib, seed, density, pad):
    rng = np.random.default_rng(seed)
    probs = rng.integers(1, 256, size=(4, 8, 3, 11), dtype=np.uint8)
    n_mb, cols = 300, 20
    enc = BoolEncoder()
    above = np.zeros((cols, 9), np.int8)
    left = np.zeros(9, np.int8)
    want = []
    a_idx = [i & 3 for i in range(16)] + [4, 5, 4, 5, 6, 7, 6, 7, 8]
    l_idx = [i >> 2 for i in range(16)] + [4, 4, 5, 5, 6, 6, 7, 7, 8]
    for m in range(n_mb):
        col = m % cols
        if col == 0:
            left[:] = 0
        has_y2 = bool(rng.integers(0, 2))
        A, L = above[col], left
        blocks, mask, eobtotal = {}, 0, -16 if has_y2 else 0
        order = ([24] if has_y2 else []) + list(range(24))
        for i in order:
            btype = 1 if i == 24 else (2 if i >= 16 else (0 if has_y2 else 3))
            first = 1 if (has_y2 and i < 16) else 0
            zz = random_block(rng, first, density)
            ctx = int(A[a_idx[i]] + L[l_idx[i]])
            eob = encode_block(enc, probs[btype], ctx, first, zz)
            A[a_idx[i]] = L[l_idx[i]] = 1 if eob > first else 0
            eobtotal += eob
            if eob > first:
                mask |= 1 << i
                raster = np.zeros(16, np.int16)
                for c in range(16):
                    raster[ZIGZAG[c]] = zz[c]
                blocks[i] = raster
        want.append((has_y2, mask, eobtotal, [blocks[i] for i in sorted(blocks)], A.copy(), L.copy()))
    data = enc.finish() + bytes(pad)
    buf = ctypes.create_string_buffer(data, len(data))
    base = ctypes.addressof(buf)
    # vp8dx_start_decode (dboolhuff.c:16-34): empty window, count -8, range 255
    bd = BoolDec(base, base + len(data), 0, -8, 255)
    above[:] = 0
    coef = np.full(25 * 16, 0x5a5a, np.int16)       # dirty arena: the reader must clear what it keeps
    mask = ctypes.c_uint32(0)
    for m, (has_y2, wmask, wtotal, wblocks, wA, wL) in enumerate(want):
        col = m % cols
        if col == 0:
            left[:] = 0
        coef[:] = 0x5a5a
        got = tokens_lib.vp8b200_decode_mb_tokens(ctypes.byref(bd), probs.ctypes.data, above[col].ctypes.data,
                                                  left.ctypes.data, int(has_y2), coef.ctypes.data, ctypes.byref(mask))
        assert (mask.value, got) == (wmask, wtotal), "macroblock %d" % m
        for k, blk in enumerate(wblocks):
            assert np.array_equal(coef[k * 16:(k + 1) * 16], blk), "macroblock %d stored block %d" % (m, k)
        assert np.array_equal(above[col], wA) and np.array_equal(left, wL), "contexts after macroblock %d" % m
    assert bd.buf <= base + len(data)


def _dump(ivf, path, ref_tokens, threads=None):
    env = dict(os.environ, VP8B200_NO_DEVICE="1", VP8B200_DUMP=path)
    env.pop("VP8B200_TOKENS", None)
    env.pop("VP8B200_PARSE_THREADS", None)
    if ref_tokens:
        env["VP8B200_TOKENS"] = "ref"
    if threads:
        env["VP8B200_PARSE_THREADS"] = str(threads)
    subprocess.run([VPXDEC_B200, "--noblit", ivf], env=env, check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL, timeout=300)
    return open(path, "rb").read()


@pytest.mark.skipif(not os.path.exists(VPXDEC_B200), reason="hostdec/_build is not built (needs the reference sources)")
@pytest.mark.parametrize("name", CASES)
def test_fused_reader_gives_the_reference_records(name):
    ivf = os.path.join(GOLD, name + ".ivf")
    with tempfile.TemporaryDirectory() as tmp:
        ref = _dump(ivf, os.path.join(tmp, "ref.rec"), True)
        new = _dump(ivf, os.path.join(tmp, "new.rec"), False)
    assert new == ref
    assert new == lzma.decompress(open(os.path.join(GOLD, name + ".rec.xz"), "rb").read())


@pytest.mark.skipif(not os.path.exists(VPXDEC_B200), reason="hostdec/_build is not built (needs the reference sources)")
@pytest.mark.parametrize("threads", [2, 3, 8])
def test_partition_parallel_parse_gives_the_serial_records(threads):
    """rows of different token partitions parsed on different threads (per-row
    arena regions, packed afterwards) must give byte-identical records to the serial parser -
    8 partitions with segmentation (w320_er8); 3 threads = uneven partition-to-thread map."""
    ivf = os.path.join(GOLD, "w320_er8.ivf")
    with tempfile.TemporaryDirectory() as tmp:
        serial = _dump(ivf, os.path.join(tmp, "t1.rec"), False, threads=1)
        par = _dump(ivf, os.path.join(tmp, "tn.rec"), False, threads=threads)
    assert par == serial
    assert par == lzma.decompress(open(os.path.join(GOLD, "w320_er8.rec.xz"), "rb").read())


def _damaged(data, mod):
    """IVF with per-frame payloads rewritten by mod(index, payload)."""
    import struct
    out, pos, i = bytearray(data[:32]), 32, 0
    while pos + 12 <= len(data):
        n = struct.unpack("<I", data[pos:pos + 4])[0]
        hdr, payload = data[pos:pos + 12], data[pos + 12:pos + 12 + n]
        pos += 12 + n
        payload = mod(i, payload)
        out += struct.pack("<I", len(payload)) + hdr[4:] + payload
        i += 1
    return bytes(out)


@pytest.mark.skipif(not os.path.exists(VPXDEC_B200), reason="hostdec/_build is not built (needs the reference sources)")
@pytest.mark.parametrize("case", ["truncated_p_frame", "truncated_key_frame", "flipped_bytes"])
def test_partition_parallel_parse_on_damaged_streams(case):
    """Truncated partitions and corrupt tokens must end the same way with the serial and the
    partition-parallel parser: same exit status, same records for the frames that decode."""
    data = open(os.path.join(GOLD, "w320_er8.ivf"), "rb").read()
    mod = {"truncated_p_frame": lambda i, f: f[:len(f) * 6 // 10] if i == 3 else f,
           "truncated_key_frame": lambda i, f: f[:len(f) // 2] if i == 0 else f,
           "flipped_bytes": lambda i, f: f[:len(f) // 2] + bytes([f[len(f) // 2] ^ 0x55]) + f[len(f) // 2 + 1:]
           if i in (2, 5) else f}[case]
    with tempfile.TemporaryDirectory() as tmp:
        ivf = os.path.join(tmp, "damaged.ivf")
        open(ivf, "wb").write(_damaged(data, mod))
        got = {}
        for threads in (1, 8):
            dump = os.path.join(tmp, "t%d.rec" % threads)
            env = dict(os.environ, VP8B200_NO_DEVICE="1", VP8B200_DUMP=dump, VP8B200_PARSE_THREADS=str(threads))
            env.pop("VP8B200_TOKENS", None)
            r = subprocess.run([VPXDEC_B200, "--noblit", ivf], env=env, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, timeout=300)
            got[threads] = (r.returncode, open(dump, "rb").read() if os.path.exists(dump) else b"")
    assert got[1] == got[8]


VPXENC = os.path.join(ROOT, "oracle", "_ref", "vpxenc")


@pytest.mark.skipif(not (os.path.exists(VPXDEC_B200) and os.path.exists(VPXENC)),
                    reason="hostdec/_build or oracle/_ref is not built (needs the reference sources)")
@pytest.mark.parametrize("token_parts,threads", [(1, 2), (2, 3), (2, 4), (3, 5)])
def test_partition_parallel_parse_other_partition_counts(token_parts, threads):
    """2, 4 and 8 token partitions encoded on the spot by the reference's vpxenc (a frame only 9
    macroblock rows high, so some threads get one row or none): the partition-parallel parser
    must reproduce the serial records."""
    import sys
    with tempfile.TemporaryDirectory() as tmp:
        y4m, ivf = os.path.join(tmp, "in.y4m"), os.path.join(tmp, "in.ivf")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_y4m.py"), "--kind", "motion", "--size",
                        "208x144", "--frames", "8", "--seed", str(40 + token_parts), "-o", y4m], check=True)
        subprocess.run([VPXENC, "--ivf", "--rt", "--cpu-used=4", "--token-parts=%d" % token_parts,
                        "--target-bitrate=400", "-o", ivf, y4m], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=300)
        serial = _dump(ivf, os.path.join(tmp, "t1.rec"), False, threads=1)
        par = _dump(ivf, os.path.join(tmp, "tn.rec"), False, threads=threads)
        ref = _dump(ivf, os.path.join(tmp, "ref.rec"), True)
    assert len(serial) > 1000 and par == serial == ref
