#!/usr/bin/env python3
"""Write the three configuration headers the reference sources expect.

Build tool shared by oracle/refbuild (the unmodified reference = the checker) and
hostdec (the reference's host parser with the B200 seams).  The reference tree ships no
pre-generated vpx_config.h / vpx_rtcd.h / vpx_version.h: its own configure +
rtcd.sh produce them.  We do not run that build system; this script writes the
equivalent of a `--target=generic-gnu --disable-multithread` configuration
(every RTCD name bound to its `_c` implementation) so that
`oracle/refbuild/Makefile` can compile the reference sources where they lie.

usage: refconfig.py <reference_root> <out_dir> [KEY=VALUE ...]   (overrides, e.g. CONFIG_B200=1)
Outputs: <out_dir>/vpx_config.h, vpx_rtcd.h, vpx_version.h, vpx_config.c
"""
import os
import re
import sys

CONFIG = {
    # arch / simd: none (generic C path = the ground truth named by the north star)
    "ARCH_ARM": 0, "ARCH_MIPS": 0, "ARCH_X86": 0, "ARCH_X86_64": 0,
    "ARCH_PPC32": 0, "ARCH_PPC64": 0,
    "HAVE_EDSP": 0, "HAVE_MEDIA": 0, "HAVE_NEON": 0, "HAVE_MIPS32": 0,
    "HAVE_MMX": 0, "HAVE_SSE": 0, "HAVE_SSE2": 0, "HAVE_SSE3": 0,
    "HAVE_SSSE3": 0, "HAVE_SSE4_1": 0, "HAVE_ALTIVEC": 0,
    "HAVE_VPX_PORTS": 1, "HAVE_STDINT_H": 1, "HAVE_ALT_TREE_LAYOUT": 0,
    "HAVE_PTHREAD_H": 1, "HAVE_SYS_MMAN_H": 1, "HAVE_DLOPEN": 0,
    "HAVE_UNISTD_H": 1,
    "CONFIG_EXTERNAL_BUILD": 0, "CONFIG_INSTALL_DOCS": 0,
    "CONFIG_INSTALL_BINS": 1, "CONFIG_INSTALL_LIBS": 1,
    "CONFIG_INSTALL_SRCS": 0, "CONFIG_DEBUG": 0, "CONFIG_GPROF": 0,
    "CONFIG_GCOV": 0, "CONFIG_RVCT": 0, "CONFIG_GCC": 1, "CONFIG_MSVS": 0,
    "CONFIG_PIC": 1, "CONFIG_BIG_ENDIAN": 0, "CONFIG_CODEC_SRCS": 0,
    "CONFIG_DEBUG_LIBS": 0, "CONFIG_FAST_UNALIGNED": 1,
    "CONFIG_MEM_MANAGER": 0, "CONFIG_MEM_TRACKER": 0, "CONFIG_MEM_CHECKS": 0,
    "CONFIG_MD5": 1, "CONFIG_DEQUANT_TOKENS": 0, "CONFIG_DC_RECON": 0,
    "CONFIG_RUNTIME_CPU_DETECT": 0, "CONFIG_POSTPROC": 0,
    "CONFIG_MULTITHREAD": 0, "CONFIG_INTERNAL_STATS": 0,
    "CONFIG_VP8_ENCODER": 1, "CONFIG_VP8_DECODER": 1, "CONFIG_VP8": 1,
    "CONFIG_ENCODERS": 1, "CONFIG_DECODERS": 1, "CONFIG_STATIC_MSVCRT": 0,
    "CONFIG_SPATIAL_RESAMPLING": 1, "CONFIG_REALTIME_ONLY": 0,
    "CONFIG_ERROR_CONCEALMENT": 0, "CONFIG_SHARED": 0, "CONFIG_STATIC": 1,
    "CONFIG_SMALL": 0, "CONFIG_OPENCL": 0, "CONFIG_POSTPROC_VISUALIZER": 0,
    "CONFIG_OS_SUPPORT": 1, "CONFIG_UNIT_TESTS": 0,
    "CONFIG_MULTI_RES_ENCODING": 0,
}


def shell_cond(line):
    """Evaluate `if [ "$CONFIG_X" = "yes" ]; then` / `!=` against CONFIG."""
    m = re.match(r'\s*if \[ "\$(\w+)" (=|!=) "yes" \]; then', line)
    if not m:
        raise SystemExit("unhandled shell line in rtcd_defs.sh: " + line)
    on = bool(CONFIG.get(m.group(1), 0))
    return on if m.group(2) == "=" else not on


def gen_rtcd(defs_path):
    out = ["#ifndef RTCD_H", "#define RTCD_H", "",
           "#ifdef RTCD_C", "#define RTCD_EXTERN", "#else",
           "#define RTCD_EXTERN extern", "#endif", ""]
    stack = []          # active if-blocks
    in_heredoc = False
    # `name_c=other_symbol` lines rename the C implementation of an RTCD name
    alias = dict(re.findall(r'^\s*(\w+_c)=(\w+)\s*$', open(defs_path).read(), re.M))
    for raw in open(defs_path):
        line = raw.rstrip("\n")
        if in_heredoc:
            if line.strip() == "EOF":
                in_heredoc = False
            else:
                out.append(line)     # forward declarations of structs
            continue
        s = line.strip()
        if s.startswith("cat <<EOF"):
            in_heredoc = True
            continue
        if s.startswith("if ["):
            stack.append(shell_cond(line))
            continue
        if s == "fi":
            stack.pop()
            continue
        if not all(stack):
            continue
        m = re.match(r'prototype\s+(.+?)\s+(\w+)\s+"(.*)"\s*$', s)
        if m:
            rtyp, name, args = m.groups()
            impl = alias.get(name + "_c", name + "_c")
            out.append("%s %s(%s);" % (rtyp, impl, args))
            out.append("#define %s %s" % (name, impl))
            out.append("")
    out += ['#include "vpx_config.h"', "", "void vpx_rtcd(void);", "",
            "#ifdef RTCD_C", "void vpx_rtcd(void)", "{", "}", "#endif",
            "#endif", ""]
    return "\n".join(out)


def main():
    ref, outdir = sys.argv[1], sys.argv[2]
    for kv in sys.argv[3:]:
        k, v = kv.split("=")
        CONFIG[k] = int(v)
    os.makedirs(outdir, exist_ok=True)
    # vpx_scale/yv12config.h includes "../vpx_config.h": the Makefile passes
    # -I<out_dir>/inc so that "<out_dir>/inc/../vpx_config.h" resolves here.
    os.makedirs(os.path.join(outdir, "inc"), exist_ok=True)
    with open(os.path.join(outdir, "vpx_config.h"), "w") as f:
        f.write("#ifndef VPX_CONFIG_H\n#define VPX_CONFIG_H\n#define RESTRICT\n")
        for k, v in CONFIG.items():
            f.write("#define %s %d\n" % (k, v))
        f.write("#endif /* VPX_CONFIG_H */\n")
    with open(os.path.join(outdir, "vpx_rtcd.h"), "w") as f:
        f.write(gen_rtcd(os.path.join(ref, "vp8/common/rtcd_defs.sh")))
    with open(os.path.join(outdir, "vpx_version.h"), "w") as f:
        f.write('#define VERSION_MAJOR  1\n#define VERSION_MINOR  0\n'
                '#define VERSION_PATCH  0\n#define VERSION_EXTRA  ""\n'
                '#define VERSION_PACKED ((VERSION_MAJOR<<16)|(VERSION_MINOR<<8)|(VERSION_PATCH))\n'
                '#define VERSION_STRING_NOSP "v1.0.0"\n'
                '#define VERSION_STRING      " v1.0.0"\n')
    with open(os.path.join(outdir, "vpx_config.c"), "w") as f:
        f.write('static const char* const cfg = "generic-gnu, no multithread (oracle/refbuild)";\n'
                'const char *vpx_codec_build_config(void) {return cfg;}\n')


if __name__ == "__main__":
    main()
