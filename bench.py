#!/usr/bin/env python3
"""bench.py - VP8 decode throughput of the B200 reconstruction path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU decoder, rank 0)

Workload (config.workload = "c5_64x1080p"): BASELINE.json configs[4] - 64 independent
1080p profile-0 streams (six-tap MC, normal loop filter; synthetic translated texture, encoded
by the reference's vpxenc) per GPU; ranks take whole streams (weak scaling, no collective).
A "step" = one frame of every stream of the rank (64 frames) reconstructed by ONE batched
launch of each kernel.

  value : frames/s with the per-frame macroblock records already resident in HBM
          (vp8b200_stage_frame + vp8b200_batch_run), CUDA events on the launching stream.
          Per step the kernels touch ~0.5 GB (64 frames x (reference + destination + records)),
          far more than the 126 MB L2, so no L2 flush is needed between steps.
  e2e   : frames/s through the reference's public API (vpx_codec_decode / vpx_codec_get_frame,
          hostdec/b200bench) from IVF bytes in host memory to vpx_image_t in host memory:
          host entropy decode + H2D of the records + kernels + D2H of every frame.
  roofline     : the loop-filter kernel (dominant), algorithmic bytes / CUDA-event time.
  cpu_baseline : the unmodified reference decoder (oracle/_ref) on the host cores, bounded sample.
"""
import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "libvpx.opencl_b200"))

STREAMS = os.path.join(ROOT, "streams")
HOSTDEC = os.path.join(ROOT, "hostdec", "_build")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
METRIC = "decode_fps_1080p_64streams_md5_exact"


def workload_config(S, F, n_clips):
    """The workload description, identical in both arms (our arm adds its own launch layout
    under "impl_config")."""
    return {"workload": "c5_64x1080p", "streams_per_gpu": S, "frames_per_clip": F, "unique_clips": n_clips,
            "resolution": "1920x1080 (coded 1920x1088)", "profile": 0, "loop_filter": "normal", "mc": "sixtap",
            "step": "one frame of each of the %d streams" % S,
            "consumer": "every visible pixel of every output frame is read (byte sum), both arms",
            "l2_policy": "inputs per step (~%.0f MB) exceed the 126 MB L2" % (S * 3 * 3428352 / 1e6)}


def assembler_found():
    """the reference's x86 SIMD build needs yasm or nasm; say whether the box has one."""
    import shutil
    return {"yasm": bool(shutil.which("yasm")), "nasm": bool(shutil.which("nasm"))}


def find_clips():
    clips = sorted(glob.glob(os.path.join(STREAMS, "c5_1080p_s*.ivf")))
    if not clips and os.path.exists(os.path.join(REFDIR, "vpxenc")):
        # normally made by __graft_entry__.build(); regenerate the same seeded streams (untimed setup)
        sys.stderr.write("bench: streams/ is empty - generating the 64 synthetic c5 streams with oracle/_ref/vpxenc\n")
        names = ["c5_1080p_s%03d" % (100 + s) for s in range(64)]
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_streams.py")] + names,
                              stdout=subprocess.DEVNULL)
        clips = sorted(glob.glob(os.path.join(STREAMS, "c5_1080p_s*.ivf")))
    if not clips:
        raise SystemExit("bench: no streams/c5_1080p_s*.ivf - run __graft_entry__.build() where oracle/_ref exists")
    return clips


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: NVML polled every few ms from a
    thread (the timed region is ~0.1 s, shorter than nvidia-smi's loop period), nvidia-smi -lms
    as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu, bus_id=None):
        self.gpu, self.bus_id, self.rows, self.p, self.nv = gpu, bus_id, [], None, None
        self.sm, self.mx, self.reasons, self.source = [], [], set(), None
        self.halt = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.bus_id:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(self.bus_id.encode())
            except pynvml.NVMLError:
                pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.gpu
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            idx = int(vis.split(",")[self.gpu])
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self):
        nv, h = self.nv
        bits = [(nv.nvmlClocksThrottleReasonHwSlowdown, 0), (nv.nvmlClocksThrottleReasonHwThermalSlowdown, 1),
                (nv.nvmlClocksThrottleReasonSwThermalSlowdown, 2), (nv.nvmlClocksThrottleReasonSwPowerCap, 3)]
        while True:                              # at least one sample even for a very short region
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.reasons.update(self.NAMES[i] for b, i in bits if r & b)
            except nv.NVMLError:
                break
            if self.halt.wait(0.004):
                break

    def start(self):
        try:
            self.nv = self._nvml_handle()
            self.mx = [float(self.nv[0].nvmlDeviceGetMaxClockInfo(self.nv[1], self.nv[0].NVML_CLOCK_SM))]
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv:
            self.halt.set()
            self.t.join(timeout=2)
        elif self.p:
            self.p.terminate()
            self.t.join(timeout=2)
            self.sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
            self.mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
            self.reasons = {self.NAMES[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"}
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None,
                "samples": len(self.sm), "source": self.source, "reasons": sorted(self.reasons)}


def bus_id(torch, dev):
    p = torch.cuda.get_device_properties(dev)
    try:
        return "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except AttributeError:
        return None


def capture_records(clips, tmpdir):
    """Untimed setup: run the host parser in record-capture mode over each clip (no pixels are
    produced in this mode; it only yields the exact C-ABI input of every frame)."""
    procs = []
    for i, c in enumerate(clips):
        out = os.path.join(tmpdir, "clip%03d.rec" % i)
        env = dict(os.environ, VP8B200_NO_DEVICE="1", VP8B200_DUMP=out)
        procs.append((out, subprocess.Popen([os.path.join(HOSTDEC, "vpxdec_b200"), "--noblit", c], env=env,
                                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
        if len(procs) % (os.cpu_count() or 8) == 0:
            for _, p in procs:
                p.wait()
    outs = []
    for out, p in procs:
        if p.wait() != 0:
            raise SystemExit("bench: record capture failed for " + out)
        outs.append(out)
    return outs


def frame_bytes_model(fr, na):
    """Algorithmic bytes per frame and kernel (DESIGN.md section 5 / SURVEY.md 8d)."""
    lvl, simple, key = int(fr.hdr["filter_level"]), int(fr.hdr["filter_type"]) != 0, int(fr.hdr["frame_type"]) == 0
    rec = 16 * fr.mb.shape[0] + 64 * fr.n_aux + 32 * fr.n_coef
    lf = 0 if lvl == 0 else (2 * na if simple else 3 * na)
    pred = 1.5 * na + rec + (0 if key else 1.5 * na)
    return {"loopfilter": lf, "pred": pred}


def run_b200(args):
    import numpy as np
    import torch
    from vp8b200 import abi, recfile, frames, shard

    rank, local_rank, world = shard.rank_info()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    clips = find_clips()
    S = args.streams
    mine = shard.streams_for_rank(rank, S, len(clips))
    uniq = sorted(set(mine))
    md5s = {u: open(clips[u][:-4] + ".md5").read().split() for u in uniq}

    # ---- setup (untimed): records of every clip this rank plays, staged in HBM -------------
    t_setup = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        recs = {}
        for u, path in zip(uniq, capture_records([clips[u] for u in uniq], tmp)):
            recs[u] = recfile.read(path)
            os.unlink(path)
    r0 = recs[uniq[0]]
    F = min(len(r.frames) for r in recs.values())
    geo = frames.Geometry(r0.coded_width, r0.coded_height)
    na = r0.coded_width * r0.coded_height
    ctxs = [abi.Context(r0.coded_width, r0.coded_height, r0.n_fb, device=local_rank) for _ in range(S)]
    staged_u = {u: [ctxs[0].stage(fr) for fr in recs[u].frames[:F]] for u in uniq}   # shared device blobs
    staged = [[staged_u[mine[s]][f] for s in range(S)] for f in range(F)]
    stream = torch.cuda.ExternalStream(ctxs[0].stream(), device=dev)
    t_setup = time.time() - t_setup

    # independent stream groups, each batched on its own CUDA stream: group A's latency-bound
    # wavefront kernels overlap group B's throughput-bound prediction kernel
    G = max(1, min(args.groups, S))
    gidx = [list(range(gi, S, G)) for gi in range(G)]
    gctx = [[ctxs[s] for s in idx] for idx in gidx]
    gstaged = [[[staged[f][s] for s in idx] for idx in gidx] for f in range(F)]
    gstreams = [torch.cuda.ExternalStream(gc[0].stream(), device=dev) for gc in gctx]

    # ctypes argument arrays are built once, not per step (the timed loop is launch-rate sensitive)
    import ctypes as C
    L = abi.lib()
    g_ca = [(C.c_void_p * len(gc))(*[c.h for c in gc]) for gc in gctx]
    g_sa = [[(C.c_void_p * len(gstaged[f][gi]))(*gstaged[f][gi]) for gi in range(G)] for f in range(F)]
    g_n = [len(gc) for gc in gctx]

    # independent streams are not in phase: group gi runs off[gi] frames ahead of group 0, so one
    # group's key frame (a latency-bound intra wavefront) overlaps the other groups' P frames
    off = [(gi * F) // G if args.stagger else 0 for gi in range(G)]

    def run_group(gi, f):
        st = L.vp8b200_batch_run(g_ca[gi], g_sa[f][gi], g_n[gi])
        if st:
            raise SystemExit("bench: vp8b200_batch_run failed: %d" % st)

    def step(i, staggered=True):
        for gi in range(G):
            run_group(gi, (i + (off[gi] if staggered else 0)) % F)

    for gi in range(G):                          # pre-roll (untimed): bring every group to its phase
        for f in range(off[gi]):
            run_group(gi, f)

    def sync_all():
        for gc in gctx:
            gc[0].sync()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: HBM-resident replay ---------------------------------------------------------
    W, K = args.warmup, args.steps
    for i in range(W):
        step(i)
    first = W % F and (F - W % F) or 0          # continue the clip so every timed step has its reference
    for i in range(W, W + first):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank, bus_id(torch, local_rank))
    sampler.start()
    sync_all()
    l0 = sum(gc[0].launch_count() for gc in gctx)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
    e0.record(gstreams[0])                       # common start: every group's stream waits for it
    for st in gstreams[1:]:
        st.wait_event(e0)
    for i in range(K):
        step(i)
    for gi in range(G):
        e1[gi].record(gstreams[gi])
    sync_all()
    barrier()
    ms = max(e0.elapsed_time(e) for e in e1)     # device time from the common start to the last group's end
    clocks = sampler.stop()
    launches = sum(gc[0].launch_count() for gc in gctx) - l0
    tot_frames, tot_s = shard.combine(dist, dev, K * S, ms / 1e3)
    value = tot_frames / tot_s

    # ---- parity (untimed): one clean pass, per-frame MD5 of fetched frames vs the reference --
    checked = 0
    check_streams = list(range(0, S, max(1, S // 4)))[:4]
    for f in range(0 if not args.skip_verify else F, F):
        step(f, staggered=False)                 # every clip restarts at its key frame
        for s in check_streams:
            fr = recs[mine[s]].frames[f]
            if fr.show_frame:
                got = frames.md5_hex(geo.i420(ctxs[s].fetch(int(fr.fb_show)), r0.display_width, r0.display_height))
                if got != md5s[mine[s]][f]:
                    raise SystemExit("bench: MD5 mismatch stream %d frame %d - result invalid" % (s, f))
                checked += 1

    # ---- per-kernel device time (events around each launch; separate pass) -------------------
    sync_all()
    ctxs[0].profile(True)
    for f in range(F):
        abi.batch_run(ctxs, staged[f])           # one group of all S streams: per-kernel times without overlap
    prof = ctxs[0].profile_read()
    ctxs[0].profile(False)
    lf_bytes = sum(frame_bytes_model(recs[mine[s]].frames[f], na)["loopfilter"] for f in range(F) for s in range(S))
    pred_bytes = sum(frame_bytes_model(recs[mine[s]].frames[f], na)["pred"] for f in range(F) for s in range(S))
    peak, peak_src = measured_peak()
    lf_ms, lf_n = prof["loopfilter"]
    kern = {k: {"ms_total": round(v[0], 3), "launches": v[1]} for k, v in prof.items()}
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    for k in kern:
        kern[k]["share"] = round(prof[k][0] / tot_ms, 4)
    border_bytes = F * S * (geo.frame_size - 1.5 * na)
    kern["border"]["GBps"] = round(border_bytes / (prof["border"][0] / 1e3) / 1e9, 1) if prof["border"][0] else None
    if kern["border"]["GBps"]:
        kern["border"]["frac"] = round(kern["border"]["GBps"] / peak, 4)
    # inter: reference read + destination write of the inter macroblocks + the records; intra: destination write
    n_inter = sum(int((recs[mine[s]].frames[f].mb["ref_frame"] != 0).sum()) for f in range(F) for s in range(S))
    n_intra = F * S * (na // 256) - n_inter
    rec_bytes = sum(16 * recs[mine[s]].frames[f].mb.shape[0] + 64 * recs[mine[s]].frames[f].n_aux + 32 * recs[mine[s]].frames[f].n_coef
                    for f in range(F) for s in range(S))
    share = n_inter / max(n_inter + n_intra, 1)
    for k, b in (("inter", n_inter * 768 + rec_bytes * share), ("intra", n_intra * 384 + rec_bytes * (1 - share))):
        if prof[k][0]:
            kern[k]["GBps"] = round(b / (prof[k][0] / 1e3) / 1e9, 1)
            kern[k]["frac"] = round(kern[k]["GBps"] / peak, 4)
    pred_ms = prof["inter"][0] + prof["intra"][0]
    kern["pred_GBps"] = round(pred_bytes / (pred_ms / 1e3) / 1e9, 1) if pred_ms else None
    achieved = lf_bytes / (lf_ms / 1e3) / 1e9 if lf_ms else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):                    # dram bytes per launch from the committed ncu capture
        traffic = json.load(open(tpath)).get("k_loopfilter", {}).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "k_loopfilter", "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "peak_source": peak_src, "bytes_per_launch": round(lf_bytes / max(lf_n, 1)),
                "ms_per_launch": round(lf_ms / max(lf_n, 1), 4), "kernels": kern}

    for c in ctxs[1:]:
        c.close()
    ctxs[0].close()

    # ---- e2e: public API, host buffers in, host frames out -----------------------------------
    e2e = None if args.skip_e2e else run_e2e(args, clips, mine, local_rank, dist, dev, shard, S)

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_refbench(clips, sample_streams=min(len(clips), os.cpu_count() or 1), repeat=1)

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        extra = north_star_extras(clips, local_rank, measured_peak()[0], recs=[recs[mine[s]] for s in range(S)], S=S)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": round(tot_s * 1e3 / K, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(S, F, len(clips)),
            "impl_config": {"stream_groups": G, "group_phase_offsets_frames": off,
                            "launches": "one batched launch per kernel, step and stream group",
                            "md5_checked_frames": checked, "setup_s": round(t_setup, 1)},
            "pixels_per_s": round(value * 1920 * 1080),
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
        }
        if extra:
            line["north_star"] = extra
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_e2e(args, clips, mine, local_rank, dist, dev, shard, S):
    # two worker threads per host core, each advancing its streams round-robin and collecting a
    # frame only after it has parsed its other streams (measured best on 16 cores: profiles/r02_summary.md)
    # ... of the cores this rank can count on: the ranks of one node share them
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = args.e2e_threads or min(S, max(2, 2 * (os.cpu_count() or 1) // world))
    repeat = max(1, args.e2e_repeat)
    env = dict(os.environ, VP8B200_DEVICE=str(local_rank), VP8B200_SYNC="block")
    cmd = [os.path.join(HOSTDEC, "b200bench"), "--threads", str(threads), "--streams", str(S),
           "--repeat", str(repeat), "--touch"] + (["--pipeline"] if args.e2e_pipeline else []) + \
          (["--delay"] if args.e2e_delay else []) + [clips[u] for u in mine]
    if dist is not None:
        dist.barrier()
    out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if out.returncode != 0:
        raise SystemExit("bench: b200bench failed: " + out.stderr[-400:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    frames_tot, wall = shard.combine(dist, dev, r["frames"], r["wall_s"])
    per_step = S / max(r["frames"], 1)
    return {"value": round(frames_tot / wall, 2), "unit": "frames/s",
            "h2d_bytes_per_step": round(r["h2d_bytes"] * per_step), "d2h_bytes_per_step": round(r["d2h_bytes"] * per_step),
            "api": "vpx_codec_decode/vpx_codec_get_frame (hostdec/b200bench)", "host_threads": r["threads"],
            "host_cores": os.cpu_count(), "frames": r["frames"], "wall_s": round(r["wall_s"], 3),
            "touch": bool(r.get("touch")), "pipeline": bool(r.get("pipeline")), "frame_delay": bool(r.get("frame_delay")),
            "checksum_per_pass": r["checksum"] // repeat if r.get("checksum") else None,
            "kernel_launches_per_frame": round(r["kernel_launches"] / max(r["frames"], 1), 3),
            "coalescer": {"batches": r.get("engine_batches"), "frames": r.get("engine_frames"),
                          "frames_per_batch": round(r["engine_frames"] / r["engine_batches"], 1) if r.get("engine_batches") else None},
            "cpu_ms_per_frame": {"decode": r["cpu_ms_per_frame_decode"], "get_frame": r["cpu_ms_per_frame_get_frame"],
                                 "blocked": r["blocked_ms_per_frame"]}}


def _b200bench(args_list, env_extra=None):
    env = dict(os.environ, VP8B200_SYNC="block")
    env.update(env_extra or {})
    out = subprocess.run([os.path.join(HOSTDEC, "b200bench")] + args_list, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True)
    if out.returncode != 0:
        raise SystemExit("bench: b200bench failed: " + out.stderr[-400:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def abi_path_rate(recs, S, local_rank, threads=16, passes=3):
    """The product's device path WITHOUT the host parser: pre-parsed records of the 64 clips go
    through the C ABI exactly as the decoder seam drives it - frame_begin, fill the pinned record
    buffers, frame_submit_show (per-device engine: batched launches, H2D of the records, D2H of
    the visible samples into a pinned host image), frame_fetch_wait - from `threads` host threads.
    Says what the API path sustains once the parser is not the limit (untimed extra)."""
    import ctypes as C
    import threading
    import numpy as np
    from vp8b200 import abi
    L = abi.lib()
    r0 = recs[0]
    F = min(len(r.frames) for r in recs)
    ctxs = [abi.Context(r0.coded_width, r0.coded_height, r0.n_fb, device=local_rank) for _ in range(S)]
    size = ctxs[0].frame_size
    ptrs = [L.vp8b200_host_alloc_on(local_rank, size) for _ in range(S)]
    outs = [np.ctypeslib.as_array((C.c_uint8 * size).from_address(p)) for p in ptrs]
    w, h = r0.display_width, r0.display_height
    bar = threading.Barrier(threads + 1)
    stats0 = (C.c_uint64 * 2)()
    g0 = abi.global_stats()

    def worker(t):
        mine = list(range(t, S, threads))
        for p in range(passes + 1):                       # pass 0 is the warm-up
            if p == 1:
                bar.wait()
            for f in range(F):
                shown = []
                for s in mine:
                    fr = recs[s % len(recs)].frames[f]
                    show = int(fr.fb_show) if fr.show_frame else -1
                    ctxs[s].submit_show(fr, show_fb=show, out=outs[s] if show >= 0 else None, display=(w, h))
                    if show >= 0:
                        shown.append(s)
                for s in shown:
                    ctxs[s].fetch_wait()
        bar.wait()

    th = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    for t in th:
        t.start()
    bar.wait()
    L.vp8b200_engine_stats(local_rank, stats0)
    g0 = abi.global_stats()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    stats1 = (C.c_uint64 * 2)()
    L.vp8b200_engine_stats(local_rank, stats1)
    g1 = abi.global_stats()
    for t in th:
        t.join()
    for c in ctxs:
        c.close()
    for p in ptrs:
        L.vp8b200_host_free(C.c_void_p(p))
    n = passes * F * S
    nb = max(1, stats1[0] - stats0[0])
    return {"fps": round(n / dt, 1), "host_threads": threads, "frames": n,
            "frames_per_batch": round((stats1[1] - stats0[1]) / nb, 1),
            "h2d_bytes_per_frame": int((g1["h2d_bytes"] - g0["h2d_bytes"]) / n),
            "d2h_bytes_per_frame": int((g1["d2h_bytes"] - g0["d2h_bytes"]) / n),
            "note": "C ABI from pre-parsed records (no bitstream parse): frame_begin + pinned record fill + "
                    "frame_submit_show + frame_fetch_wait, python threads"}


def north_star_extras(clips, local_rank, peak, recs=None, S=64):
    """Untimed extras the north star asks for next to the headline (rank 0, N = 1): single-stream
    1080p / 2160p frames/s with the records resident and through the public API (blocking call
    order and the opt-in frame-delay mode), the host parser's own ceiling on this box, and the
    x86-assembler probe for the reference's SIMD build."""
    from vp8b200 import abi, recfile
    dev_env = {"VP8B200_DEVICE": str(local_rank)}
    out = {"host_cores": os.cpu_count(), "assembler_on_box": assembler_found()}
    streams = [("1080p", clips[0])]
    c4 = os.path.join(STREAMS, "c4_2160p.ivf")
    if os.path.exists(c4):
        streams.append(("2160p_8partitions", c4))
    with tempfile.TemporaryDirectory() as tmp:
        for name, ivf in streams:
            rec = recfile.read(capture_records([ivf], tmp)[0])
            F = len(rec.frames)
            ctx = abi.Context(rec.coded_width, rec.coded_height, rec.n_fb, device=local_rank)
            staged = [ctx.stage(fr) for fr in rec.frames]
            ctx.profile(True)
            best = None
            for _ in range(3):
                for f in range(F):
                    abi.batch_run([ctx], [staged[f]])
                prof = ctx.profile_read()
                ms = sum(v[0] for v in prof.values())
                if best is None or ms < best[0]:
                    best = (ms, prof)
            ctx.close()
            na = rec.coded_width * rec.coded_height
            lf_bytes = sum(frame_bytes_model(fr, na)["loopfilter"] for fr in rec.frames)
            row = {"frames": F, "resident_fps": round(F / (best[0] * 1e-3), 1),
                   "resident_pixels_per_s": round(F / (best[0] * 1e-3) * rec.display_width * rec.display_height),
                   "kernel_ms_per_frame": {k: round(v[0] / F, 4) for k, v in best[1].items()},
                   "loopfilter_frac_of_hbm": round(lf_bytes / (best[1]["loopfilter"][0] * 1e-3) / 1e9 / peak, 4) if best[1]["loopfilter"][0] else None}
            r = _b200bench(["--threads", "1", "--streams", "1", "--repeat", "3", "--touch", ivf], dev_env)
            row["e2e_fps_blocking_call_order"] = round(r["fps"], 1)
            r = _b200bench(["--threads", "1", "--streams", "1", "--repeat", "3", "--touch", "--delay", ivf], dev_env)
            row["e2e_fps_frame_delay_mode"] = round(r["fps"], 1)
            r = _b200bench(["--threads", "1", "--streams", "1", "--repeat", "3", ivf], {"VP8B200_NO_DEVICE": "1"})
            row["host_parse_only_fps"] = round(r["fps"], 1)
            out[name] = row
    cores = os.cpu_count() or 1
    r = _b200bench(["--threads", str(cores), "--streams", "64", "--repeat", "2"] + clips[:64], {"VP8B200_NO_DEVICE": "1"})
    out["host_parse_only_fps_all_cores_64_streams"] = round(r["fps"], 1)
    if recs:
        try:
            out["abi_path_64_streams"] = abi_path_rate(recs, S, local_rank, threads=min(16, cores))
        except Exception as e:                            # an extra must never cost the headline
            out["abi_path_64_streams"] = {"error": str(e)[:200]}
    return out


def run_refbench(clips, sample_streams, repeat, procs=None):
    procs = procs or os.cpu_count() or 1
    cmd = [os.path.join(REFDIR, "refbench"), "--procs", str(procs), "--repeat", str(repeat), "--touch"] + clips[:sample_streams]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if out.returncode != 0:
        raise SystemExit("bench: refbench failed: " + out.stderr[-400:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    return {"value": round(r["fps"], 2), "unit": "frames/s", "cores": r["procs"], "kind": "reference",
            "sample": "%d 1080p clips x %d pass(es) = %d frames, unmodified reference generic-C decoder, "
                      "one process per core, every output pixel read" % (sample_streams, repeat, r["frames"]),
            "build": "generic C; x86 SIMD build needs yasm/nasm: %s" % json.dumps(assembler_found()),
            "wall_s": round(r["wall_s"], 3)}


def run_reference(args):
    """The reference's own CPU implementation of the path, all host cores, same config/metric:
    the UNMODIFIED reference decoder (oracle/_ref, generic C) decodes exactly --steps frames of
    each of the 64 streams, one process per core, and reads every output pixel like our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = find_clips()
    S = args.streams
    use = [clips[i % len(clips)] for i in range(S)]
    F = len(open(use[0][:-4] + ".md5").read().split())
    procs = os.cpu_count() or 1
    base = [os.path.join(REFDIR, "refbench"), "--procs", str(procs), "--touch"]
    if args.warmup:
        subprocess.run(base + ["--frames", str(args.warmup)] + use[:procs], stdout=subprocess.DEVNULL)
    out = subprocess.run(base + ["--frames", str(args.steps)] + use, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if out.returncode != 0:
        raise SystemExit("bench: refbench failed: " + out.stderr[-400:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    steps = r["frames"] / S
    cpu = {"value": round(r["fps"], 2), "unit": "frames/s", "cores": r["procs"], "kind": "reference",
           "build": "generic C (--target=generic-gnu equivalent, -O3), no assembler on the box: %s" % json.dumps(assembler_found()),
           "sample": "%d streams x %d frames, one process per core, every output pixel read" % (S, args.steps)}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(r["fps"], 2), "unit": "frames/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": int(round(steps)), "warmup": args.warmup,
        "ms_per_step": round(r["wall_s"] * 1e3 / steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(S, F, len(clips)),
        "impl_config": {"note": "CPU only: host cores do not scale with --gpus; rank 0 runs, others idle",
                        "checksum": r.get("checksum")},
        "cpu_baseline": cpu,
        "e2e": {"value": round(r["fps"], 2), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=90)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=64, help="independent streams per GPU")
    ap.add_argument("--groups", type=int, default=2, help="stream groups batched on separate CUDA streams")
    ap.add_argument("--stagger", type=int, default=1,
                    help="1: stream group g plays g*F/G frames ahead (streams out of phase); 0: all streams on the same frame")
    ap.add_argument("--e2e-threads", type=int, default=0)
    ap.add_argument("--e2e-repeat", type=int, default=4)
    ap.add_argument("--e2e-pipeline", type=int, default=1,
                    help="1: a worker collects a frame only after parsing its other streams (decode only queues device work)")
    ap.add_argument("--e2e-delay", type=int, default=0, help="1: decoder frame-delay mode (VP8B200_FRAME_DELAY)")
    ap.add_argument("--no-extra", action="store_true", help="skip the untimed north-star extras (single-stream numbers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: kernels only")
    ap.add_argument("--skip-verify", action="store_true", help="profiling runs: no MD5 pass")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
