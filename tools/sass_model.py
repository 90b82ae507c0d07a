#!/usr/bin/env python3
"""Static single-warp issue model of a kernel's SASS (no GPU needed).

Decodes the scheduling control bits of every sm_100a instruction (stall count, yield, write /
read barrier slot, wait mask: bits 105..121 of the 128-bit encoding) from `cuobjdump -sass` and
prints the listing with them, plus per-range totals: number of instructions, sum of the stall
fields (= the cycles ONE warp needs to issue the range when no scoreboard wait fires) and the
instruction mix by pipe.  Used to compare variants of the chain-bound kernels (loop filter,
intra) before spending GPU time: the dependent-issue time of one macroblock iteration is what
bounds them.

usage: tools/sass_model.py <lib.so> <kernel-name-substring> [--range LO HI] [--list]
       (LO / HI are hex instruction addresses as printed in the listing)
"""
import argparse
import collections
import re
import subprocess
import sys

ALU = {"IADD3", "LOP3", "SHF", "PRMT", "VABSDIFF", "VABSDIFF4", "VIMNMX", "VIMNMX3", "VIADD", "VIADDMNMX",
       "ISETP", "SEL", "IMNMX", "LEA", "IABS", "POPC", "FLO", "BREV", "SGXT", "BMSK", "PLOP3", "MOV", "CS2R",
       "FMNMX", "FSETP", "FSEL", "SHL", "SHR", "IADD", "LOP", "P2R", "R2P", "VOTE", "VOTEU", "IDP", "IDP4A"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
LSU = {"LDS", "STS", "LDG", "STG", "LDGSTS", "LDSM", "LD", "ST", "ATOMG", "ATOMS", "RED", "LDC", "LDL", "STL",
       "LDGDEPBAR", "DEPBAR", "SHFL", "MATCH", "REDUX"}
CBU = {"BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "BREAK", "RET", "CALL", "NANOSLEEP", "YIELD", "JMP"}


def pipe_of(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base in CBU:
        return "cbu"
    if base.startswith("U") or base in ("S2UR", "R2UR"):
        return "uniform"
    if base in ALU:
        return "alu"
    return "other"


def parse(lib, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    out, on, pend = [], False, None
    for line in txt.splitlines():
        if "Function :" in line:
            on = kernel in line
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m:
            pend = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", line)
        if m and pend:
            hi = int(m.group(1), 16)
            addr, text, lo = pend
            pend = None
            t = text.split()
            pred = t[0] if t[0].startswith("@") else ""
            op = t[1] if pred else t[0]
            out.append({"addr": addr, "text": text, "op": op, "pred": pred,
                        "stall": (hi >> 41) & 0xf, "yield": (hi >> 45) & 1, "wbar": (hi >> 46) & 7,
                        "rbar": (hi >> 49) & 7, "wait": (hi >> 52) & 0x3f})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("kernel")
    ap.add_argument("--range", nargs=2, action="append", default=[])
    ap.add_argument("--list", action="store_true")
    a = ap.parse_args()
    ins = parse(a.lib, a.kernel)
    if not ins:
        raise SystemExit("no such kernel")
    if a.list:
        for i in ins:
            w = "".join(str(b) if (i["wait"] >> b) & 1 else "-" for b in range(6))
            print("%05x  s%-2d %s w%s r%s wait[%s]  %s" % (i["addr"], i["stall"], "Y" if i["yield"] else " ",
                  i["wbar"] if i["wbar"] < 6 else "-", i["rbar"] if i["rbar"] < 6 else "-", w, i["text"]))
    ranges = [(int(lo, 16), int(hi, 16)) for lo, hi in a.range] or [(ins[0]["addr"], ins[-1]["addr"])]
    for lo, hi in ranges:
        sel = [i for i in ins if lo <= i["addr"] <= hi]
        mix = collections.Counter(pipe_of(i["op"]) for i in sel)
        ops = collections.Counter(i["op"].split(".")[0] for i in sel)
        print("range %05x..%05x: %d instructions, stall sum %d cycles, waits on scoreboard %d" %
              (lo, hi, len(sel), sum(i["stall"] for i in sel), sum(1 for i in sel if i["wait"])))
        print("  pipes:", dict(mix))
        print("  top ops:", ops.most_common(14))


if __name__ == "__main__":
    main()
