#!/usr/bin/env python
"""bench.py -- throughput of the dense-stereo hot path (MDE/s = W*H*D disparity evaluations per second, and fps).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1]

A "step" is one stereo frame through the whole path (AD-Census cost for 2 views -> CBCA x2 iterations -> 8-path
SGM -> WTA -> LR check -> region vote x2 -> interpolation x2 -> 3x3 median).  The default workload is BASELINE.json's
metric configuration, 1920x1080 with D=256 (configs[2]/[4]), synthetic texture-warped pairs.

  value  : whole-job MDE/s with the frame already resident in HBM (sm_pipeline_run_device), CUDA events.
  e2e    : the same through the host-buffer C-ABI call sm_pipeline_run: pinned host images in (H2D), left
           disparity map out (D2H), copies inside the timed region.
  roofline / stages : per-stage device time measured live (CUDA events the pipeline records on its stream
           during the timed steps) against the algorithmic HBM bytes of SURVEY.md section 8(d).
  cpu_baseline : the compiled reference (oracle/_ref/libsmref.so; else the oracle port) on bounded bands of the
           same workload, on this box's host cores (N=1, rank 0 only).

N>1 (torchrun, one process per GPU): frames are independent, each rank runs its own K frames (weak scaling), no
data-path collective; torch.distributed is used only for the barrier and the max-over-ranks time.

--impl reference : the reference's own CPU code for the path -- its hot-path function bodies cut from
stereoMatching.{h,cpp} and compiled against a cv::Mat stand-in (oracle/_ref/libsmref.so, oracle/build_ref_sm.py);
one instance per host thread (the reference is single-threaded), a bounded band per instance per step.  Falls back
to the oracle port when that library is absent or the workload uses NL aggregation.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, D, paths, kind)
    "c1": (450, 375, 64, 4, "random_dot"),
    "c2": (1280, 720, 128, 8, "texture_warped"),
    "c3": (1920, 1080, 256, 8, "texture_warped"),
    "c4": (640, 480, 64, 4, "texture_warped"),       # NL non-local MST aggregation instead of CBCA
    "c3cg": (1920, 1080, 256, 8, "texture_warped"),  # c3 with the cost main_.cpp:15 compiles in: censusGrad
    "c3main": (1920, 1080, 256, 8, "texture_warped"),  # the reference's main() as compiled: censusGrad + SolveAll(PY_LEV 1, 0.3)
    "c3py3": (1920, 1080, 256, 8, "texture_warped"),   # c3 with the caller's cross-scale step over a 3-level pyramid
    "c5": (1920, 1080, 256, 8, "texture_warped"),      # BASELINE configs[4]: the c3 frame as a 64-frame stream, strong scaling
    "c3cen": (1920, 1080, 256, 8, "texture_warped"),   # costcalculation "Census", no aggregation: the native uint16 path (b = 2)
}
AGGREGATION = {"c1": 1, "c2": 1, "c3": 1, "c4": 2, "c3cg": 1, "c3main": 1, "c3py3": 1, "c5": 1, "c3cen": 0}
STREAM_FRAMES = 64     # BASELINE configs[4]: "batched 1080p D=256 synthetic stereo stream, frame-parallel at 1/2/4/8 B200"
COSTCALC = {"c3cg": 1, "c3main": 1, "c3cen": 2}   # 0 = AD-Census (BASELINE configs), 1 = censusGrad, 2 = Census (uint16 volumes)
PYRAMID = {"c3main": (1, 0.3), "c3py3": (3, 0.3)}   # (PY_LEV, REG_LAMBDA) of main_.cpp:132, 157; absent: no SolveAll


def workload_desc(name):
    W, H, D, P, kind = WORKLOADS[name]
    agg = {0: "no aggregation", 1: "CBCA(2 it, intersected arms)", 2: "NL(MST tree filter, sigma 0.1, left view)"}[AGGREGATION[name]]
    cost = {0: "AD-Census(71-bit)", 1: "censusGrad(71-bit census + arm-weighted x/y gradient)",
            2: "Census(71-bit Hamming, uint16 volumes)"}[COSTCALC.get(name, 0)]
    if name in PYRAMID:
        agg += f"+SolveAll({PYRAMID[name][0]} level(s), lambda {PYRAMID[name][1]})"
    return (f"{name}: {W}x{H} D={D} {cost}+{agg}+{P}-path SGM+WTA+LRC+"
            f"regionVote x2+properIpol x2+median3, 2 views, {'uint16' if COSTCALC.get(name, 0) == 2 else 'fp32'} volumes, "
            f"synthetic {kind} pairs")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def l2_note(W, H, D, b):
    """How the timed steps relate to the 126 MB L2 (timing rule: inputs larger than L2, or a flush)."""
    vol = W * H * D * b / 1e6
    if vol > 126:
        return f"inputs larger than L2: each pass streams {vol / 1e3:.1f} GB volumes (L2 = 126 MB)"
    return (f"a frame touches 3-4 volumes of {vol:.0f} MB plus their scratch (more than the 126 MB L2 in total, every pass "
            f"writes a volume the next one reads); no separate flush between steps: the only data that survives from one "
            f"step to the next is the {8 * W * H / 1e6:.1f} MB of input images")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- CPU arms
def cpu_band(name, rows):
    """A `rows`-high band of the workload (full width and D): MDE/s is size-normalised, so the band's rate is the
    workload's rate up to the vertical-arm / vertical-path extent."""
    from mystereomatching_b200 import synth
    W, H, D, P, kind = WORKLOADS[name]
    rows = min(rows, H)
    return synth.make_pair(rows, W, D, kind, seed=1000), rows


def run_oracle(pair, name, threads):
    from oracle import pyoracle as po
    W, H, D, P, kind = WORKLOADS[name]
    po.lib().orc_set_threads(threads)
    lv, lam = PYRAMID.get(name, (1, -1.0))
    op = po.default_params(D, paths=P, aggregation=AGGREGATION[name], costcalc=COSTCALC.get(name, 0), pyr_levels=lv,
                           cross_lambda=lam)
    t0 = time.perf_counter()
    _, _, _, ms = po.pipeline(pair["bgrL"], pair["bgrR"], pair["grayL"], pair["grayR"], op)
    return time.perf_counter() - t0, ms


def host_threads():
    from oracle import pyoracle as po
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return n if po.lib().orc_has_openmp() else 1


def have_compiled_reference(name):
    """oracle/_ref/libsmref.so = the reference's own stereoMatching.{h,cpp} function bodies (oracle/build_ref_sm.py);
    for c4 (aggregation == "NL") also its StereoMatching::NL + NL/NLCCA.cpp over the compiled NL/ sources (libqxref.so)."""
    from oracle import pyoracle as po
    L = po.smref_lib()
    if L is None:
        return False
    return AGGREGATION[name] == 1 or hasattr(L, "smref_pipeline_nl")


def reference_threads(name, rows):
    """The reference is single-threaded (every `omp parallel for` in it is commented out, stereoMatching.h:596, 648,
    1661), so 'all the host threads it can use' = one independent instance per core, each on its own band --
    the same frame-level partition the GPU side uses.  Capped by memory: an instance holds ~88 B per (pixel, d)
    (vm x2, four ADCensusCal temporaries, three cbca temporaries, L[8], HVL_INTERSECTION x2)."""
    W, H, D, P, kind = WORKLOADS[name]
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per = rows * W * D * 88
    try:
        import psutil
        n = max(1, min(n, int(0.5 * psutil.virtual_memory().available / per)))
    except Exception:
        n = min(n, 8)
    return min(n, 32)


def run_reference(name, rows, threads, seed0=1000):
    """`threads` compiled-reference StereoMatching instances, each running the whole default chain on its own
    W x rows band.  Returns wall seconds."""
    from oracle import pyoracle as po
    from mystereomatching_b200 import synth
    W, H, D, P, kind = WORKLOADS[name]
    pairs = [synth.make_pair(rows, W, D, kind, seed=seed0 + i) for i in range(threads)]

    def work(p):
        r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
        if AGGREGATION[name] == 2:
            r.pipeline_nl(P)
        elif name in PYRAMID:
            r.pipeline_pyr(PYRAMID[name][0], PYRAMID[name][1], P, 2, COSTCALC.get(name, 0))
        else:
            r.pipeline(P, 2, costcalc=COSTCALC.get(name, 0))   # ctypes releases the GIL for the duration of the call
        r.close()

    ts = [threading.Thread(target=work, args=(p,)) for p in pairs]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return time.perf_counter() - t0


def cpu_baseline(name, budget_rows=48):
    W, H, D, P, kind = WORKLOADS[name]
    if have_compiled_reference(name):
        rows = H if H * W * D <= 64 * 2 ** 20 else min(H, 96)   # whole frames when small; else tall enough for the vertical arms (<= 34 rows either way)
        nt = reference_threads(name, rows)
        mde = W * rows * D / 1e6
        t1 = run_reference(name, rows, 1)
        tn = run_reference(name, rows, nt) if nt > 1 else t1
        return {"value": nt * mde / tn, "unit": "MDE/s", "cores": nt, "kind": "reference",
                "value_1thread": mde / t1,
                "sample": f"the reference's own code (oracle/_ref/libsmref.so): whole chain on {W}x{rows} bands (full "
                          f"width, D={D}); one band on 1 thread {t1:.1f} s (the reference's own threading), "
                          f"{nt} bands on {nt} threads {tn:.1f} s"}
    pair, rows = cpu_band(name, budget_rows)
    mde = W * rows * D / 1e6
    nt = host_threads()
    t1, _ = run_oracle(pair, name, 1)
    tn, _ = run_oracle(pair, name, nt) if nt > 1 else (t1, None)
    return {"value": mde / tn, "unit": "MDE/s", "cores": nt, "kind": "port",
            "value_1thread": mde / t1,
            "sample": f"one {W}x{rows} band (full width, D={D}) of the workload through the whole oracle chain; "
                      f"{t1:.1f} s at 1 thread (the reference's own threading), {tn:.1f} s at {nt} threads"}


def parity_vs_oracle(name, frame, disp_gpu):
    """The measured frame through the oracle restatement (all host threads): share of identical pixels of the refined
    left map and bad-2 of both.  Bounded: the frame is cut to the rows the host memory holds (the GPU map is then
    compared on a re-run of that band by the caller)."""
    from oracle import pyoracle as po
    from mystereomatching_b200 import synth
    W, H, D, P, kind = WORKLOADS[name]
    po.lib().orc_set_threads(host_threads())
    lv, lam = PYRAMID.get(name, (1, -1.0))
    op = po.default_params(D, paths=P, aggregation=AGGREGATION[name], costcalc=COSTCALC.get(name, 0), pyr_levels=lv,
                           cross_lambda=lam)
    t0 = time.perf_counter()
    ref, _, _, _ = po.pipeline(frame["bgrL"], frame["bgrR"], frame["grayL"], frame["grayR"], op)
    dt = time.perf_counter() - t0
    return {"vs": "oracle (oracle/stereo_oracle.cpp, bit-equal to the compiled reference on tests/golden/sm_ref.npz)",
            "pct_identical": round(100.0 * float((ref == disp_gpu).mean()), 5),
            "bad2_nonocc_pct_oracle": round(synth.bad_k(ref, frame["gt"], frame["nonocc"], 2), 3),
            "frame": f"{W}x{ref.shape[0]} D={D}, seed 1000 (the measured frame 0 of rank 0), whole refined left map",
            "oracle_seconds": round(dt, 1), "oracle_threads": host_threads()}


def oracle_fits(name):
    W, H, D, P, kind = WORKLOADS[name]
    try:
        import psutil
        return psutil.virtual_memory().available > W * H * D * 4 * (8 + P) * 1.3
    except Exception:
        return False


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    W, H, D, P, kind = WORKLOADS[name]
    total = args.steps + args.warmup
    compiled = have_compiled_reference(name)
    if compiled:
        # band height of a timed step: tall enough that the vertical arms (<= 34 rows either way) and the vertical SGM
        # paths are not cut short the way a 12-row band cuts them; whole frames for the small workloads.  Sized so that
        # the default K = 10 stays within a few minutes on 16 host threads (~0.17 s per 1920-wide D=256 row and step).
        rows = H if H * W * D <= 64 * 2 ** 20 else min(H, 96 if args.steps <= 12 else (64 if args.steps <= 20 else 32))
        nt = reference_threads(name, rows)
        for i in range(args.warmup):   # warm-up steps only touch code and caches: a short band each
            run_reference(name, min(rows, 24), nt, 1000 + 100 * i)
        dt = 0.0
        for i in range(args.steps):
            dt += run_reference(name, rows, nt, 5000 + 100 * i)
        mde = nt * W * rows * D / 1e6
        kindb = "reference"
        sample = (f"each step = {nt} instances of the reference's own StereoMatching code (oracle/_ref/libsmref.so), one "
                  f"per host thread, each running the whole chain on its own {W}x{rows} band (full width, D={D})")
        note = ("the reference's hot-path function bodies, cut from stereoMatching.{h,cpp} and compiled against a "
                "cv::Mat stand-in (oracle/build_ref_sm.py); the reference itself is single-threaded")
    else:
        rows = 48 if total <= 14 else (24 if total <= 40 else 12)
        pair, rows = cpu_band(name, rows)
        nt = host_threads()
        for _ in range(args.warmup):
            run_oracle(pair, name, nt)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run_oracle(pair, name, nt)
        dt = time.perf_counter() - t0
        mde = W * rows * D / 1e6
        kindb = "port"
        sample = f"each step = one {W}x{rows} band (full width, D={D}) through the whole oracle chain, {nt} threads"
        note = ("no compiled reference for this workload here (oracle/_ref/libsmref.so absent, or NL aggregation): "
                "this arm times the oracle port of the reference's algorithm (oracle/stereo_oracle.cpp)")
    val = mde * args.steps / dt
    out = {"impl": "reference", "metric": "MDE/s (W*H*D disparity evaluations per second), whole path",
           "value": val, "unit": "MDE/s", "fps_equiv": val * 1e6 / (W * H * D), "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_desc(name), "parallelism": f"host CPU, {nt} threads"},
           "cpu_baseline": {"value": val, "unit": "MDE/s", "cores": nt, "kind": kindb, "sample": sample},
           "e2e": {"value": val, "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "note": note}
    emit(out)
    return 0


# ------------------------------------------------------------------------------------------------- GPU arm
def sgm_is_grouped(name):
    """Shapes the grouped row sweeps (k_sgm_group) accept: mirrors smi_sgm_group's own check (sgm_group.cu)."""
    W, H, D, P, kind = WORKLOADS[name]
    nb = min(148, W // 4)
    return P == 8 and D % 4 == 0 and 64 < D <= 256 and nb >= 1 and -(-W // nb) <= 14


def stage_bytes(name):
    """Algorithmic HBM bytes per LAUNCH and launches per frame for each volume stage (b = 4)."""
    W, H, D, P, kind = WORKLOADS[name]
    V = W * H * D
    b = 2 if COSTCALC.get(name, 0) == 2 else 4
    if COSTCALC.get(name, 0) == 2:     # native uint16 path: Hamming volume write, then SGM on uint16 volumes
        if P == 8 and D in (128, 256) and -(-W // min(148, W // 4)) <= 14:
            # per view: grouped sweep UP (C in, S out = 2 V b), grouped sweep DOWN (3 V b), paths 2 and 3 (3 V b each; the right
            # view's last path only feeds the fused WTA: 2 V b) -> 21 V b per frame in 8 launches
            sg = {"bytes_per_launch": 21.0 / 8 * V * b, "launches": 8, "kernel": "k_sgm_group_u16 + k_sgm_path_u16_h"}
        else:                          # 8 single-path sweeps per view
            sg = {"bytes_per_launch": (3 * P - 1 - 0.5) * V * b / P, "launches": 2 * P, "kernel": "k_sgm_path_u16"}
        return {"cost": {"bytes_per_launch": V * b, "launches": 2, "kernel": "k_cost<HAMMING_U16>"},
                "sgm": sg,
                "wta": {"bytes_per_launch": 0, "launches": 2, "kernel": "fused into k_sgm_path_u16 (mode 2 / 3)"}}
    agg = {"bytes_per_launch": 2 * V * b, "launches": 8, "kernel": "k_cbca_pass"}
    if AGGREGATION[name] == 2:   # k_nl = 4 volume passes; MST build + rooting + two tree sweeps
        agg = {"bytes_per_launch": 4 * V * b, "launches": 1, "kernel": "nl (Boruvka MST + Euler-tour rooting + k_tf_cta_fast)"}
    st = {
        "cost": {"bytes_per_launch": V * b, "launches": 2,
                 "kernel": "k_cost_grad<FUSED>" if COSTCALC.get(name, 0) else "k_cost<ADCENSUS>"},
        "aggregation": agg,
        # gen_dispFromVm is fused into the last SGM path of each view (no separate read of the summed volume)
        "wta": {"bytes_per_launch": 0, "launches": 2, "kernel": "fused into k_sgm_path (mode 2)"},
    }
    if sgm_is_grouped(name):
        # per view: sweep UP writes S (C + S = 2 V b), sweep DOWN read-modify-writes it (3 V b); then paths 2 and 3
        # (C, S in, S out = 3 V b each; the fused WTA of path 3 still leaves the summed volume behind as vm)
        # (D <= 128: both views share a launch, two CTAs per SM: 2 launches of twice the bytes)
        nl = 2 if D <= 128 else 4
        st["sgm_group"] = {"bytes_per_launch": 10.0 / nl * V * b, "launches": nl, "kernel": "k_sgm_group"}
        # paths 2 and 3 of both views: C, S in, S out = 3 V b each, except the right view's last path, whose finished sum
        # only feeds the fused WTA and is not stored (keep_right_volume = 0): 2 V b
        st["sgm_path"] = {"bytes_per_launch": 11.0 / 4 * V * b, "launches": 4, "kernel": "k_sgm_path"}
    else:
        # first path writes S (2 V b), the others read-modify-write it (3 V b)
        # <= 1024 scan lines per sweep: the cp.async-staged any-direction kernel (sgm.cu, SM_SGM_SMALL_LINES)
        small = max(W, H) + (min(W, H) - 1 if P == 8 else 0) <= 1024
        st["sgm"] = {"bytes_per_launch": (3 * P - 1 - 0.5) * V * b / P, "launches": 2 * P,
                     "kernel": "k_sgm_path_s" if small else "k_sgm_path"}
    return st


def stream_frame(name, i, base_cache):
    """Frame i of the synthetic stream: 8 base pairs (seeds 1000 .. 1007, ~3 s of numpy each at 1080p), frame i = base
    i % 8 rolled down by 16 * (i // 8) rows -- a vertical roll keeps the row-wise stereo geometry, and all 64 frames differ."""
    import numpy as np
    from mystereomatching_b200 import synth
    W, H, D, P, kind = WORKLOADS[name]
    b = i % 8
    if b not in base_cache:
        base_cache[b] = synth.make_pair(H, W, D, kind, seed=1000 + b)
    shift = 16 * (i // 8)
    pr = base_cache[b]
    return {k: np.ascontiguousarray(np.roll(pr[k], shift, axis=0)) for k in ("bgrL", "bgrR", "grayL", "grayR")}


def run_stream_bench(name, params, local, rank, world, barrier, max_over_ranks, n_frames=STREAM_FRAMES):
    """BASELINE configs[4] through the C ABI's frame stream (sm_stream_*): this rank's frames (i mod world == rank) are
    submitted from pinned host buffers, every left map comes back to the host; wall clock from the first submit to the last
    retired frame, max over ranks.  Whole job = n_frames frames whatever N is (strong scaling)."""
    import numpy as np
    import torch
    from mystereomatching_b200 import capi
    W, H, D, P, kind = WORKLOADS[name]
    mine = list(range(rank, n_frames, world))
    cache = {}
    host = []
    for i in mine:
        fr = stream_frame(name, i, cache)
        host.append({k: torch.from_numpy(v).pin_memory() for k, v in fr.items()})
    outs = [torch.empty((H, W), dtype=torch.int16).pin_memory() for _ in mine]
    st = capi.Stream([local], H, W, params, queue_depth=4)

    def go(idx):
        tk = []
        for j in idx:
            h = host[j]
            tk.append(st.submit(h["bgrL"].numpy(), h["bgrR"].numpy(), h["grayL"].numpy(), h["grayR"].numpy(), outs[j].numpy()))
        st.drain()
        return tk

    go(range(min(3, len(mine))))           # warm-up: first-use allocations, clocks
    l0 = st.launches()
    barrier()
    t0 = time.perf_counter()
    go(range(len(mine)))
    dt = time.perf_counter() - t0
    barrier()
    wall = max_over_ranks(dt)
    launches = st.launches() - l0
    chk = int(sum(int(o.numpy()[H // 2, W // 2]) for o in outs))
    first = outs[0].numpy().copy() if mine else None
    st.close()
    mde = n_frames * W * H * D / 1e6
    return {"workload": f"c5: {n_frames}-frame {W}x{H} D={D} stream (frame i = base pair 1000 + i%8 rolled by 16*(i//8) rows), "
                        f"frame i -> rank i mod N, the c3 pipeline per frame",
            "api": "sm_stream_create / sm_stream_submit / sm_stream_wait (C ABI): one worker thread + sm_pipeline per GPU, "
                   "H2D of frame i+1 and D2H of frame i-1 overlapped with the compute of frame i; pinned host buffers",
            "frames": n_frames, "frames_this_rank": len(mine), "wall_ms": 1e3 * wall, "fps": n_frames / wall,
            "value": mde / wall, "unit": "MDE/s", "scaling": "strong", "timing": "wall clock, max over ranks",
            "h2d_bytes_per_frame": 8 * W * H, "d2h_bytes_per_frame": 2 * W * H, "gpu_launches_this_rank": launches,
            "checksum": chk}, first


def main_ours(args):
    import numpy as np
    import torch
    from mystereomatching_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # NCCL's own log lines (NCCL_DEBUG as the launcher set it) go to fd 1, which QuietStdout points at stderr
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    name = args.workload
    W, H, D, P, kind = WORKLOADS[name]
    ctx = capi.Ctx(local)
    lv, lam = PYRAMID.get(name, (1, -1.0))
    params = capi.default_params(D - 1, sgm_paths=P, aggregation=AGGREGATION[name],
                                 costcalculation=COSTCALC.get(name, 0), pyramidLevels=lv, crossScaleLambda=lam)
    pl = capi.Pipeline(ctx, H, W, params)

    # frames of this rank: frame i of the stream uses seed 1000+i, frame i -> rank i mod N
    n_distinct = 2
    frames = []
    for j in range(n_distinct):
        pr = synth.make_pair(H, W, D, kind, seed=1000 + rank + j * world)
        pinned = {}
        for k in ("bgrL", "bgrR", "grayL", "grayR"):
            t = torch.from_numpy(pr[k]).pin_memory()
            pinned[k] = t
        frames.append((pr, pinned))
    out_pinned = torch.empty((H, W), dtype=torch.int16).pin_memory()
    out_np = out_pinned.numpy()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def host_args(j):
        p = frames[j % n_distinct][1]
        return [p[k].numpy() for k in ("bgrL", "bgrR", "grayL", "grayR")]

    # ---- device-resident throughput (value) + live per-stage timing
    pl.upload(*host_args(0))
    for _ in range(args.warmup):
        pl.run_device()
    # clocks and power state settle over a few hundred ms of load: keep warming (untimed) until ~0.5 s has been spent
    extra_warm = 0
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.5 and extra_warm < 40:
        pl.run_device()
        torch.cuda.synchronize()
        extra_warm += 1
    pl.enable_timing(True)
    stage_acc = {}
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pl.run_device()
        for k, v in pl.stage_ms().items():      # syncs this frame; the gap before the next launch is ~10 us
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launches() - l0
    pl.enable_timing(False)

    # ---- end to end through the host-buffer call (pinned host in, host disparity out)
    for j in range(max(1, args.warmup)):
        pl.run(*host_args(j), out=out_np)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record()
    chk = 0
    for j in range(args.steps):
        pl.run(*host_args(j), out=out_np)
        chk += int(out_np[H // 2, W // 2])       # the result is read on the host every step
    f1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), wall_ms if dist is None else f0.elapsed_time(f1)))
    clocks = sampler.stop() if rank == 0 else None

    # sanity of the measured frames: quality against the synthetic ground truth (not part of the timing)
    disp = pl.run(*host_args(0)).copy()
    bad2 = synth.bad_k(disp, frames[0][0]["gt"], frames[0][0]["nonocc"], 2)

    stream_res = None
    if name in ("c3", "c5") and not args.no_stream:
        stream_res, first_map = run_stream_bench(name, params, local, rank, world, barrier, max_over_ranks)
        if rank == 0 and first_map is not None:
            ref0 = pl.run(*[stream_frame(name, 0, {})[k] for k in ("bgrL", "bgrR", "grayL", "grayR")])
            stream_res["frame0_equals_sm_pipeline_run"] = bool((ref0 == first_map).all())

    mde_frame = W * H * D / 1e6
    value = world * args.steps * mde_frame / (dev_ms / 1e3)
    e2e_val = world * args.steps * mde_frame / (e2e_ms / 1e3)
    peak, peak_src = peaks()
    sb = stage_bytes(name)
    stages = {}
    for k, info in sb.items():
        ms = stage_acc.get(k, 0.0) / args.steps
        per_launch_ms = ms / info["launches"]
        gbs = info["bytes_per_launch"] / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        stages[k] = {"ms_per_frame": round(ms, 4), "launches": info["launches"], "kernel": info["kernel"],
                     "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
    for k in ("census", "arms", "refine", "total") + (("sgm",) if "sgm" not in sb else ()) + (("aggregation",) if "aggregation" not in sb else ()):
        stages[k] = {"ms_per_frame": round(stage_acc.get(k, 0.0) / args.steps, 4)}
    dom = max(sb, key=lambda k: stages[k]["ms_per_frame"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(name, {}).get(sb[dom]["kernel"])
        except Exception:
            traffic = None
    total_alg = sum(v["bytes_per_launch"] * v["launches"] for v in sb.values())
    roofline = {"bound": "hbm", "kernel": sb[dom]["kernel"], "achieved": stages[dom]["GBps"], "peak": peak,
                "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sb[dom]["bytes_per_launch"],
                "whole_frame": {"algorithmic_GB": round(total_alg / 1e9, 2),
                                "achieved_GBps": round(total_alg / 1e9 / (dev_ms / 1e3 / args.steps), 1),
                                "frac": round(total_alg / 1e9 / (dev_ms / 1e3 / args.steps) / peak, 4)}}

    if rank == 0:
        out = {"metric": "MDE/s (W*H*D disparity evaluations per second), whole path",
               "value": value, "unit": "MDE/s", "fps": value * 1e6 / (W * H * D), "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u16" if COSTCALC.get(name, 0) == 2 else "f32", "data": "synthetic",
               "config": {"workload": workload_desc(name), "frames_per_step_per_gpu": 1,
                          "extra_untimed_warmup_steps": extra_warm,
                          "parallelism": f"frame-parallel x{world}, no collective",
                          "l2": l2_note(W, H, D, 2 if COSTCALC.get(name, 0) == 2 else 4)},
               "e2e": {"value": e2e_val, "unit": "MDE/s", "fps": e2e_val * 1e6 / (W * H * D),
                       "ms_per_step": e2e_ms / args.steps,
                       "h2d_bytes_per_step": 8 * W * H, "d2h_bytes_per_step": 2 * W * H},
               "gpu_launches": launches, "roofline": roofline, "stages": stages, "clocks": clocks,
               "quality": {"bad2_nonocc_pct": round(bad2, 3), "checksum": chk}}
        if stream_res is not None:
            out["stream"] = stream_res
            if name == "c5":     # the stream IS the workload: whole-job throughput over the 64 frames, strong scaling
                out["value"], out["fps"], out["scaling"] = stream_res["value"], stream_res["fps"], "strong"
                out["ms_per_step"] = stream_res["wall_ms"] / stream_res["frames"]
                out["e2e"] = {"value": stream_res["value"], "unit": "MDE/s", "fps": stream_res["fps"],
                              "ms_per_step": stream_res["wall_ms"] / stream_res["frames"],
                              "h2d_bytes_per_step": 8 * W * H, "d2h_bytes_per_step": 2 * W * H}
                out["config"]["note"] = ("value = e2e = the 64-frame stream through sm_stream_* with host buffers (wall clock, max "
                                         "over ranks); the device-resident per-stage numbers below are the c3 frame's")
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(name)
            if oracle_fits(name):
                out["parity"] = parity_vs_oracle(name, frames[0][0], disp)
            else:
                out["parity"] = {"vs": "oracle", "pct_identical": None, "note": "host memory too small for a whole-frame oracle run"}
        emit(out)
    pl.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


class QuietStdout:
    """stdout carries exactly ONE JSON line.  Native libraries (NCCL's version banner, the reference's printf
    chatter) write to file descriptor 1 directly, so for the duration of the run fd 1 is pointed at stderr and the
    JSON line is written to the saved descriptor at the end."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def close(self):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


OUT = None


def emit(obj):
    line = json.dumps(obj)
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line)


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-stream", action="store_true", help="skip the 64-frame stream leg (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    OUT = QuietStdout()
    try:
        return main_reference(args) if args.impl == "reference" else main_ours(args)
    finally:
        OUT.close()


if __name__ == "__main__":
    sys.exit(main())
