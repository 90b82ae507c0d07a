"""YV12 frame-buffer geometry (the reference layout, vpx_scale/generic/yv12config.c:55-110)
and the visible-area hash `vpxdec --md5 --i420` computes (vpxdec.c:1093-1115)."""
import hashlib

import numpy as np

BORDER = 32


class Geometry:
    def __init__(self, coded_w, coded_h):
        assert coded_w % 16 == 0 and coded_h % 16 == 0
        self.w, self.h = coded_w, coded_h
        self.y_stride = ((coded_w + 2 * BORDER) + 31) & ~31
        self.uv_stride = self.y_stride >> 1
        self.yplane = (coded_h + 2 * BORDER) * self.y_stride
        self.uvplane = ((coded_h >> 1) + BORDER) * self.uv_stride
        self.frame_size = self.yplane + 2 * self.uvplane
        self.y_off = BORDER * self.y_stride + BORDER
        self.u_off = self.yplane + (BORDER // 2) * self.uv_stride + BORDER // 2
        self.v_off = self.yplane + self.uvplane + (BORDER // 2) * self.uv_stride + BORDER // 2

    def planes(self, buf):
        """(Y, U, V) views of the coded area inside a whole-allocation uint8 array."""
        buf = np.asarray(buf, np.uint8).reshape(-1)
        ys, us = self.y_stride, self.uv_stride
        y = np.lib.stride_tricks.as_strided(buf[self.y_off:], (self.h, self.w), (ys, 1))
        u = np.lib.stride_tricks.as_strided(buf[self.u_off:], (self.h // 2, self.w // 2), (us, 1))
        v = np.lib.stride_tricks.as_strided(buf[self.v_off:], (self.h // 2, self.w // 2), (us, 1))
        return y, u, v

    def i420(self, buf, disp_w, disp_h):
        """The bytes vpxdec writes/hashes for one frame: d_w x d_h luma, (d+1)/2 chroma."""
        y, u, v = self.planes(buf)
        cw, ch = (disp_w + 1) // 2, (disp_h + 1) // 2
        return (np.ascontiguousarray(y[:disp_h, :disp_w]).tobytes() +
                np.ascontiguousarray(u[:ch, :cw]).tobytes() +
                np.ascontiguousarray(v[:ch, :cw]).tobytes())

    def defined_mask(self):
        """Boolean mask over the allocation of the bytes the reference defines after border
        extension: everything except per-row padding beyond width + 2*border."""
        m = np.zeros(self.frame_size, bool)
        rows_y = self.h + 2 * BORDER
        my = m[:self.yplane].reshape(rows_y, self.y_stride)
        my[:, :self.w + 2 * BORDER] = True
        for off in (self.yplane, self.yplane + self.uvplane):
            mc = m[off:off + self.uvplane].reshape((self.h >> 1) + BORDER, self.uv_stride)
            mc[:, :(self.w >> 1) + BORDER] = True
        return m


def md5_hex(b):
    return hashlib.md5(b).hexdigest()
