"""Reader / writer for ".rec" files (include/vp8b200_recfile.h): the per-frame macroblock
records the host parser hands to the C ABI, captured to disk."""
import struct

import numpy as np

REC_MAGIC = 0x52385056        # "VP8R"
FRAME_MAGIC = 0x314D5246      # "FRM1"

MB_DTYPE = np.dtype([("y_mode", "u1"), ("uv_mode", "u1"), ("ref_frame", "u1"), ("flags", "u1"),
                     ("mv_row", "<i2"), ("mv_col", "<i2"), ("coef_mask", "<u4"), ("coef_off", "<u4")])
assert MB_DTYPE.itemsize == 16

# vp8b200_frame_hdr, 76 bytes
HDR_DTYPE = np.dtype([
    ("frame_type", "u1"), ("use_bilinear_mc", "u1"), ("full_pixel", "u1"), ("filter_type", "u1"),
    ("filter_level", "u1"), ("sharpness_level", "u1"), ("segmentation_enabled", "u1"),
    ("segment_abs_delta", "u1"), ("mode_ref_lf_delta_enabled", "u1"), ("fb_new", "u1"),
    ("fb_last", "u1"), ("fb_golden", "u1"), ("fb_altref", "u1"), ("reserved", "u1", (3,)),
    ("segment_lf", "i1", (4,)), ("ref_lf_deltas", "i1", (4,)), ("mode_lf_deltas", "i1", (4,)),
    ("dequant", "<i2", (4, 3, 2))])
assert HDR_DTYPE.itemsize == 76

FILE_HDR = struct.Struct("<8I")
FRAME_HDR = struct.Struct("<4I2B2x")          # followed by the 76-byte vp8b200_frame_hdr

MBF_SKIP = 0x04
MBF_CLAMP = 0x08


class Frame:
    __slots__ = ("hdr", "mb", "aux", "coef", "show_frame", "fb_show")

    def __init__(self, hdr, mb, aux, coef, show_frame, fb_show):
        self.hdr, self.mb, self.aux, self.coef = hdr, mb, aux, coef
        self.show_frame, self.fb_show = show_frame, fb_show

    @property
    def n_aux(self):
        return self.aux.shape[0]

    @property
    def n_coef(self):
        return self.coef.shape[0]


class RecFile:
    def __init__(self, display, coded, n_fb, frames):
        self.display_width, self.display_height = display
        self.coded_width, self.coded_height = coded
        self.n_fb = n_fb
        self.frames = frames


def read(path, max_frames=None):
    with open(path, "rb") as f:
        data = f.read()
    return parse(data, max_frames)


def parse(data, max_frames=None):
    magic, ver, dw, dh, cw, ch, n_fb, _ = FILE_HDR.unpack_from(data, 0)
    if magic != REC_MAGIC:
        raise ValueError("not a .rec file")
    pos = FILE_HDR.size
    frames = []
    while pos < len(data) and (max_frames is None or len(frames) < max_frames):
        fmagic, n_mb, n_aux, n_coef, show, fb_show = FRAME_HDR.unpack_from(data, pos)
        if fmagic != FRAME_MAGIC:
            raise ValueError("bad frame magic at %d" % pos)
        pos += FRAME_HDR.size
        hdr = np.frombuffer(data, HDR_DTYPE, 1, pos)[0].copy()
        pos += HDR_DTYPE.itemsize
        mb = np.frombuffer(data, MB_DTYPE, n_mb, pos).copy()
        pos += 16 * n_mb
        aux = np.frombuffer(data, np.uint8, 64 * n_aux, pos).reshape(n_aux, 64).copy()
        pos += 64 * n_aux
        coef = np.frombuffer(data, "<i2", 16 * n_coef, pos).reshape(n_coef, 16).copy()
        pos += 32 * n_coef
        frames.append(Frame(hdr, mb, aux, coef, show, fb_show))
    return RecFile((dw, dh), (cw, ch), n_fb, frames)


def write(path, rec):
    with open(path, "wb") as f:
        f.write(FILE_HDR.pack(REC_MAGIC, 1, rec.display_width, rec.display_height,
                              rec.coded_width, rec.coded_height, rec.n_fb, 0))
        for fr in rec.frames:
            f.write(FRAME_HDR.pack(FRAME_MAGIC, fr.mb.shape[0], fr.n_aux, fr.n_coef,
                                   int(fr.show_frame), int(fr.fb_show)))
            f.write(np.asarray(fr.hdr, HDR_DTYPE).tobytes())
            f.write(np.ascontiguousarray(fr.mb).tobytes())
            f.write(np.ascontiguousarray(fr.aux).tobytes())
            f.write(np.ascontiguousarray(fr.coef, "<i2").tobytes())
