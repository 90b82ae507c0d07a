"""Seeded synthetic macroblock records (valid by construction) that exercise every mode,
edge case and clamp of the reconstruction path far more densely than encoded streams do."""
import numpy as np

from vp8b200 import recfile
from vp8b200.recfile import HDR_DTYPE, MB_DTYPE, Frame


def random_frame(rng, mb_cols, mb_rows, key=False, bilinear=False, full_pixel=False,
                 filter_type=0, filter_level=None, sharpness=None, segmentation=None,
                 p_intra=0.15, p_split=0.2, p_skip=0.3, coef_density=0.4, big_coefs=False,
                 fbs=(0, 1, 2, 3)):
    n_mb = mb_cols * mb_rows
    hdr = np.zeros((), HDR_DTYPE)
    hdr["frame_type"] = 0 if key else 1
    hdr["use_bilinear_mc"] = int(bilinear)
    hdr["full_pixel"] = int(full_pixel)
    hdr["filter_type"] = filter_type
    hdr["filter_level"] = rng.integers(0, 64) if filter_level is None else filter_level
    hdr["sharpness_level"] = rng.integers(0, 8) if sharpness is None else sharpness
    seg = bool(rng.integers(0, 2)) if segmentation is None else segmentation
    hdr["segmentation_enabled"] = int(seg)
    hdr["segment_abs_delta"] = int(rng.integers(0, 2))
    hdr["mode_ref_lf_delta_enabled"] = int(rng.integers(0, 2))
    hdr["fb_new"], hdr["fb_last"], hdr["fb_golden"], hdr["fb_altref"] = fbs
    if hdr["segment_abs_delta"]:
        hdr["segment_lf"] = rng.integers(0, 64, 4)
    else:
        hdr["segment_lf"] = rng.integers(-40, 40, 4)
    hdr["ref_lf_deltas"] = rng.integers(-20, 20, 4)
    hdr["mode_lf_deltas"] = rng.integers(-20, 20, 4)
    dq = np.zeros((4, 3, 2), np.int16)
    for s in range(4):
        dq[s, 0] = rng.integers(4, 158, 2)
        dq[s, 1] = rng.integers(8, 480, 2)
        dq[s, 2] = rng.integers(4, 158, 2)
    if not seg:
        dq[:] = dq[0]
    hdr["dequant"] = dq

    mb = np.zeros(n_mb, MB_DTYPE)
    aux, coefs = [], []
    for i in range(n_mb):
        row, col = divmod(i, mb_cols)
        intra = key or rng.random() < p_intra
        m = mb[i]
        flags = int(rng.integers(0, 4))
        if intra:
            m["y_mode"] = rng.integers(0, 5)
            m["uv_mode"] = rng.integers(0, 4)
            m["ref_frame"] = 0
            if m["y_mode"] == 4:
                a = np.zeros(64, np.uint8)
                a[:16] = rng.integers(0, 10, 16)
                m["mv_row"], m["mv_col"] = np.array([len(aux)], "<u4").view("<i2")
                aux.append(a)
        else:
            m["ref_frame"] = rng.integers(1, 4)
            m["uv_mode"] = 0
            split = rng.random() < p_split
            clamp = rng.random() < 0.3
            # legal range when the clamp flag is off: what the parser guarantees
            lo_c, hi_c = -(col * 16 + 16) * 8, ((mb_cols - 1 - col) * 16 + 16) * 8
            lo_r, hi_r = -(row * 16 + 16) * 8, ((mb_rows - 1 - row) * 16 + 16) * 8

            def one_mv():
                if clamp:
                    r = int(rng.integers(lo_r - 400, hi_r + 400))
                    c = int(rng.integers(lo_c - 400, hi_c + 400))
                else:
                    span = 8 * 24 if rng.random() < 0.7 else 10 ** 6
                    r = int(rng.integers(max(lo_r, -span), min(hi_r, span) + 1))
                    c = int(rng.integers(max(lo_c, -span), min(hi_c, span) + 1))
                r, c = r & ~1, c & ~1                    # luma MVs are even (decodemv.c:110-114)
                if rng.random() < 0.25:
                    r &= ~7
                if rng.random() < 0.25:
                    c &= ~7
                return max(min(r, 32767), -32768), max(min(c, 32767), -32768)
            if clamp:
                flags |= recfile.MBF_CLAMP
            if split:
                m["y_mode"] = 9
                a = np.zeros(32, "<i2")
                kind = rng.integers(0, 4)
                base = [one_mv() for _ in range(16)]
                for b in range(16):
                    src = {0: b, 1: (b // 8) * 8, 2: (b % 4) // 2 * 2, 3: (b // 8) * 8 + (b % 4) // 2 * 2}[int(kind)]
                    a[2 * b], a[2 * b + 1] = base[src]
                m["mv_row"], m["mv_col"] = np.array([len(aux)], "<u4").view("<i2")
                aux.append(a.view(np.uint8))
            else:
                m["y_mode"] = rng.choice([5, 6, 7, 8])
                m["mv_row"], m["mv_col"] = (0, 0) if m["y_mode"] == 7 and rng.random() < 0.5 else one_mv()
        skip = rng.random() < p_skip
        m["coef_off"] = len(coefs)
        if skip:
            flags |= recfile.MBF_SKIP
        else:
            has_y2 = m["y_mode"] not in (4, 9)
            mask = 0
            for b in range(25):
                if b == 24 and not has_y2:
                    continue
                if rng.random() < coef_density:
                    mask |= 1 << b
                    c = np.zeros(16, np.int16)
                    nz = rng.integers(1, 17)
                    idx = rng.choice(16, nz, replace=False)
                    lim = 32767 if (big_coefs and rng.random() < 0.2) else 60
                    c[idx] = rng.integers(-lim, lim + 1, nz)
                    if b < 16 and has_y2:
                        c[0] = 0                         # DC position belongs to the WHT
                    coefs.append(c)
            m["coef_mask"] = mask
        m["flags"] = flags
    aux_a = np.stack(aux) if aux else np.zeros((0, 64), np.uint8)
    coef_a = np.stack(coefs) if coefs else np.zeros((0, 16), np.int16)
    return Frame(hdr, mb, aux_a, coef_a, 1, int(hdr["fb_new"]))


def random_buffers(rng, frame_size, n_fb):
    return [rng.integers(0, 256, frame_size, dtype=np.uint8) for _ in range(n_fb)]
