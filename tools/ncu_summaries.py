#!/usr/bin/env python3
"""Turns the ncu reports of tools/round_end_measure.sh (r02_inter / r02_lf / r02_intra .ncu-rep in
DIR) into the short text summaries kept under profiles/.
usage: tools/ncu_summaries.py gpurun_out/final profiles"""
import csv, subprocess, sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_barriers", "gpu__time_duration.sum",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return [dict(zip(rows[0], r)) for r in rows[2:]], dict(zip(rows[0], rows[1]))


def fmt(x):
    try:
        return "%.3f" % float(x) if "." in x else x
    except (TypeError, ValueError):
        return str(x)


def stalls(v, n=8):
    st = sorted(((float(x), k) for k, x in v.items()
                 if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k), reverse=True)
    return ", ".join("%s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], x) for x, k in st[:n])


def table(o, cols, units):
    o.write("%-74s" % "metric" + "".join(" %22s" % name for name, _ in cols) + "\n")
    for k in KEYS:
        o.write("%-74s" % (k + " [" + units.get(k, "") + "]") + "".join(" %22s" % fmt(v.get(k, "n/a")) for _, v in cols) + "\n")


def roof(v):
    inst = float(v["smsp__inst_executed.sum"])
    return inst, inst / (4 * 148 * 1.965e9) * 1e6


def main(src, dst):
    hdr = "ncu --set full --clock-control none --import-source on; times under ncu are cold-cache and serialised,\nthe CUDA-event times are in r02_summary.md and on the bench line\n"
    rows, units = load(src + "/r02_inter.ncu-rep")
    with open(dst + "/r02_inter_summary.txt", "w") as o:
        o.write("k_inter16, 64 x 1080p P frame per launch - " + hdr)
        o.write("algorithmic bytes per launch: read 1.5*Na*64 = 200.5 MB + write 200.5 MB + records ~ 12 MB\n")
        o.write("earlier variants of the round (same capture command, reports not kept): 8 lanes per MB with direct loads 296 us,\n"
                "205.7 M instructions, 72 registers, IPC 2.53; cp.async-staged 358 us, 213.5 M, 88 registers, IPC 2.13,\n"
                "shared-memory wavefronts 65.5 % of peak; round 1: 277.5 M instructions, IPC 3.1\n\n")
        table(o, [("shipped: split warps", rows[0])], units)
        inst, us = roof(rows[0])
        o.write("\nwarp stall reasons (per issue): " + stalls(rows[0]) + "\n")
        o.write("instruction-issue roof: %.1f M warp instructions / (4 x 148 x 1.965 GHz) = %.0f us at 100 %% issue\n" % (inst / 1e6, us))
    rows, units = load(src + "/r02_intra.ncu-rep")
    with open(dst + "/r02_intra_summary.txt", "w") as o:
        o.write("k_intra, 64 x 1080p per launch - " + hdr)
        o.write("launch 1 = KEY frame (8160 intra macroblocks per stream, all B_PRED), launch 2 = first P frame\n")
        o.write("round 1 (r01_intra_v8_summary.txt): key frame 1.683 ms, 1,292 M warp instructions, IPC 2.66, DRAM 436 + 267 MB\n\n")
        table(o, [("key frame", rows[0]), ("P frame", rows[1])], units)
        inst, us = roof(rows[0])
        o.write("\nwarp stall reasons, key frame (per issue): " + stalls(rows[0]) + "\n")
        o.write("key frame: %.0f warp instructions per macroblock; instruction-issue roof %.0f us at 100 %% issue\n" % (inst / 522240, us))
        o.write("DRAM: the frame itself is 200 MB of the writes; the rest is the 128-byte message block per macroblock (67 MB\n"
                "written, read back by up to four neighbours, mostly from L2) and, on the read side, the key frame's\n"
                "coefficient records (dense on a key frame) and macroblock records\n")
    rows, units = load(src + "/r02_lf.ncu-rep")
    with open(dst + "/r02_lf_summary.txt", "w") as o:
        o.write("k_loopfilter, 64 x 1080p P frame per launch (step 4 of the clip) - " + hdr)
        o.write("algorithmic bytes per launch: 3*Na*64 = 400.2 MB (read + write of Y, U, V); CUDA-event time 0.486-0.493 ms =\n"
                "812-824 GB/s = 12.4-12.6 % of the measured 6549 GB/s\n")
        o.write("rejected packed 16x2 variant (r02_lf_packed_summary.txt): 568 us, 122.9 M instructions, 164 registers, IPC 0.93, issue 25 %,\n"
                "6.8 warps per SM, stall barrier 1.20 / long_scoreboard 1.51\n\n")
        table(o, [("shipped", rows[0])], units)
        inst, us = roof(rows[0])
        o.write("\nwarp stall reasons (per issue): " + stalls(rows[0], 10) + "\n")
        o.write("instruction-issue roof: %.1f M warp instructions / (4 x 148 x 1.965 GHz) = %.0f us at 100 %% issue; the integer ALU pipe\n"
                "re-issues every 2 cycles (tools/sass_model.py), so ~%.0f us is the practical floor of this instruction mix\n" % (inst / 1e6, us, us / 0.65))
        o.write("""
where the instructions go (ncu source page joined with nvdisasm line info):
  81 % of the macroblock iterations (422 k of 522 k) take the no-inner-edge path (mb_skip_coeff set, no
  SPLITMV / B_PRED): only the two macroblock edges are filtered.
  filter arithmetic: left MB edge 13.4 %, top MB edge 14.9 %, inner edges 6.6 % + 7.2 % of all warp
  instructions = 42 %; the other 58 % are the per-macroblock frame of the row chain: parameter
  broadcast, cp.async ring (address + predicate arithmetic 6 %), pack / unpack (8 %), tile moves,
  hand-off send / receive, stores, loop and reconvergence bookkeeping (BSSY / BSYNC / BRA 7 %).
  33 % of the stall samples sit on the named-barrier wait for the row above: rows can never run
  ahead of the row above, so every hiccup of an upper row propagates down the frame.
""")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
