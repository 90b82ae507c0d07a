#!/usr/bin/env python3
"""Copy the reference's HOST decoder sources into a scratch directory and cut the B200 seams.

The product keeps the reference's bitstream parser (north star: "bitstream parsing stays in
C on the host").  Nothing from the reference is committed to this repository: this script
copies the needed files from <reference_root> into <out_dir> at build time and applies
anchor-based edits (each anchor must match exactly once, otherwise the build fails loudly,
so a drifted reference cannot be patched silently).  The edits are the reference-side
binding INTEGRATION.md documents; every inserted line calls into hostdec/vp8b200_seam.c.

usage: apply_seams.py <reference_root> <out_dir>
"""
import os
import shutil
import sys

# directories / files of the reference the host decoder build needs
COPY = ["vp8/common", "vp8/decoder", "vp8/vp8_dx_iface.c", "vpx", "vpx_mem", "vpx_ports",
        "vpx_scale", "vpxdec.c", "md5_utils.c", "md5_utils.h", "args.c", "args.h",
        "tools_common.c", "tools_common.h", "nestegg"]

INC = '#include "vp8b200_seam.h"\n'


def find_one(lines, needle, start=0):
    hits = [i for i in range(start, len(lines)) if needle in lines[i]]
    if len(hits) != 1:
        raise SystemExit("apply_seams: anchor %r matched %d times" % (needle, len(hits)))
    return hits[0]


def patch(path, fn):
    with open(path) as f:
        lines = f.readlines()
    fn(lines)
    with open(path, "w") as f:
        f.writelines(lines)


def p_onyxd_int(l):
    i = find_one(l, "} VP8D_COMP;")
    l.insert(i, "    void *b200_seam;   /* hostdec/vp8b200_seam.c state */\n")


def p_decodframe(l):
    i = find_one(l, '#include "onyxd_int.h"')
    l.insert(i + 1, INC)
    # S2: frame begin replaces the 127/129 edge setup (device applies the rule itself)
    i = find_one(l, "vp8_setup_intra_recon(&pc->yv12_fb[pc->new_fb_idx]);")
    l[i] = "        vp8b200_seam_frame_begin(pbi);\n"
    # S3a: tokens are read straight into the coefficient arena
    i = find_one(l, "eobtotal = vp8_decode_mb_tokens(pbi, xd);")
    l[i] = l[i].replace("vp8_decode_mb_tokens", "vp8b200_seam_decode_tokens")
    # S3: per-MB record replaces prediction + residual
    i = find_one(l, "/* do prediction */")
    l.insert(i, "    vp8b200_seam_record_mb(pbi, xd, mb_idx);\n    return;\n")
    # partition-parallel token parse: rows may run on several threads, so the
    # row function clears the left context it was given, waits for the row above before a
    # macroblock and publishes the macroblock afterwards; the row loop asks the seam first
    i = find_one(l, "vpx_memset(&pc->left_context, 0, sizeof(pc->left_context));")
    l[i] = "    vpx_memset(xd->left_context, 0, sizeof(*xd->left_context));\n"
    i = find_one(l, "decode_macroblock(pbi, xd, mb_row * pc->mb_cols  + mb_col);")
    l.insert(i + 1, "        vp8b200_seam_mb_done(mb_row, mb_col);\n")
    l.insert(i, "        vp8b200_seam_mb_wait(mb_row, mb_col);\n")
    i = find_one(l, "static unsigned int read_partition_size(const unsigned char *cx_size)")
    l.insert(i, "static void decode_mb_row_b200(void *pbi, int mb_row, void *xd)\n{\n"
                "    decode_mb_row((VP8D_COMP *)pbi, &((VP8D_COMP *)pbi)->common, mb_row, (MACROBLOCKD *)xd);\n}\n\n")
    i = find_one(l, "decode_mb_row(pbi, pc, mb_row, xd);")
    j = i
    while "for (mb_row = 0; mb_row < pc->mb_rows; mb_row++)" not in l[j]:
        j -= 1
    assert i - j < 16
    l[j] = l[j].replace("mb_row = 0;", "mb_row = vp8b200_seam_decode_rows(pbi, xd, decode_mb_row_b200) ? pc->mb_rows : 0;")
    # per-row 4-pixel extension: device rule, drop the host call (spans several lines)
    i = find_one(l, "vp8_extend_mb_row(")
    j = i
    while ");" not in l[j]:
        j += 1
    l.insert(j + 1, "#endif\n")
    l.insert(i, "#if 0 /* vp8b200: done on the device */\n")


def p_onyxd_if(l):
    i = find_one(l, '#include "onyxd_int.h"')
    l.insert(i + 1, INC)
    # S5: loop filter + border extension -> one asynchronous device submit
    a = find_one(l, "if(cm->filter_level)")
    b = find_one(l, "vp8_yv12_extend_frame_borders_ptr(cm->frame_to_show);")
    assert a < b
    l[a:b + 1] = ["        vp8b200_seam_frame_submit(pbi);\n"]
    # get_raw_frame: only QUEUE the device->host copy of the shown buffer; the
    # wait sits in vp8_get_frame.  Frame-delay mode is handled by the seam at the entry.
    i = find_one(l, "*sd = *pbi->common.frame_to_show;")
    l[i] = "        if (vp8b200_seam_show(pbi, sd)) return -1;\n"
    i = find_one(l, "int ret = -1;")
    l.insert(i + 1, "    { int r_ = vp8b200_seam_get_raw_frame(pbi, sd, time_stamp, time_end_stamp); if (r_ != 1) return r_; }\n")
    # decode(NULL, 0) is the flush request in frame-delay mode, not a lost frame
    i = find_one(l, "if (!pbi->ec_active &&")
    l.insert(i, "    if (pbi->num_fragments <= 1 && pbi->fragment_sizes[0] == 0 && vp8b200_seam_flush(pbi))\n"
                "    {\n        pbi->num_fragments = 0;\n        return 0;\n    }\n")
    # VP8_COPY_REFERENCE / VP8_SET_REFERENCE: the device owns the pixels
    i = find_one(l, "vp8_yv12_copy_frame_ptr(&cm->yv12_fb[ref_fb_idx], sd);")
    l[i] = "    {\n        if (vp8b200_seam_sync_fb(pbi, ref_fb_idx)) return pbi->common.error.error_code;\n" + l[i] + "    }\n"
    i = find_one(l, "vp8_yv12_copy_frame_ptr(sd, &cm->yv12_fb[*ref_fb_ptr]);")
    l.insert(i + 1, "        if (vp8b200_seam_upload_fb(pbi, *ref_fb_ptr)) return pbi->common.error.error_code;\n")
    i = find_one(l, "vp8_remove_common(&pbi->common);")
    l.insert(i, "    vp8b200_seam_destroy(pbi);\n")
    # missing-frame path: device-side copy next to the host copy
    i = find_one(l, "vp8_yv12_copy_frame_ptr(&cm->yv12_fb[prev_idx],")
    j = i
    while ");" not in l[j]:
        j += 1
    l.insert(j + 1, "            vp8b200_seam_copy_fb(pbi, cm->lst_fb_idx, prev_idx);\n")


def p_dx_iface(l):
    i = find_one(l, '#include "decoder/onyxd_int.h"')
    l.insert(i + 1, INC)
    # vpx_codec_get_frame is where the caller waits for the device
    i = find_one(l, "img = &ctx->img;")
    l.insert(i, "            if (vp8b200_seam_wait((struct VP8D_COMP *)ctx->pbi))\n"
                "            {\n                ctx->img_avail = 0;\n                return NULL;\n            }\n")
    # a device failure while queueing the copy is a decode error, not "no frame"
    i = find_one(l, "ctx->img_avail = 1;")
    assert l[i + 1].strip() == "}"
    l.insert(i + 2, "        if (!res)\n            res = update_error_state(ctx, &((struct VP8D_COMP *)ctx->pbi)->common.error);\n")


def p_yv12config(l):
    i = find_one(l, '#include "vpx_mem/vpx_mem.h"')
    l.insert(i + 1, INC)
    i = find_one(l, "vpx_memalign(32, ybf->frame_size)")
    l[i] = l[i].replace("vpx_memalign(32, ybf->frame_size)", "vp8b200_seam_alloc(ybf->frame_size)")
    i = find_one(l, "vpx_free(ybf->buffer_alloc);")
    l[i] = l[i].replace("vpx_free(ybf->buffer_alloc)", "vp8b200_seam_free(ybf->buffer_alloc)")


def main():
    ref, out = sys.argv[1], sys.argv[2]
    if os.path.exists(out):
        shutil.rmtree(out)
    os.makedirs(out)
    for item in COPY:
        s, d = os.path.join(ref, item), os.path.join(out, item)
        if os.path.isdir(s):
            shutil.copytree(s, d)
        else:
            os.makedirs(os.path.dirname(d) or ".", exist_ok=True)
            shutil.copy(s, d)
    for root, _, files in os.walk(out):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    patch(os.path.join(out, "vp8/decoder/onyxd_int.h"), p_onyxd_int)
    patch(os.path.join(out, "vp8/decoder/decodframe.c"), p_decodframe)
    patch(os.path.join(out, "vp8/decoder/onyxd_if.c"), p_onyxd_if)
    patch(os.path.join(out, "vpx_scale/generic/yv12config.c"), p_yv12config)
    patch(os.path.join(out, "vp8/vp8_dx_iface.c"), p_dx_iface)
    print("apply_seams: patched host decoder in", out)


if __name__ == "__main__":
    main()
