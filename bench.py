#!/usr/bin/env python
"""Benchmark of the develop hot path (BASELINE.json metric: Mpix/s of AHD develop on 24 MP RGGB).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline (BASELINE config 2: QualityDemosaic.Best, postprocess_stages=1, WB + camera->linear sRGB, float32 out).
A "step" develops FRAMES distinct synthetic 6000x4000 14-bit RGGB frames per GPU, `loops` times over, where `loops` is
chosen after the warm-up so that the K timed steps last at least 2 s (reported in `config`).
  value     whole-job Mpix/s with the mosaics already resident in HBM (CUDA events, max over ranks);
  e2e       the same metric through the public batch API (pysp_b200.pipeline.FramePipeline) with HOST buffers: pinned
            H2D of every mosaic and D2H of every result inside the timed region, >= 20 steps; next to it the rate of
            the same copies with no kernels (`ceiling_mpix_s`: what the box's host<->device path allows) and
            `e2e_variants` for the narrower outputs (lin_f16, srgb_u8);
  roofline  dominant kernel: algorithmic bytes per launch / its mean device time (CUDA events recorded by the library
            around each launch on the launching stream) against the measured HBM copy peak; `traffic` / `issue` are
            constants from one ncu capture (their `*_source` says which);
  workloads device-resident Mpix/s of the other BASELINE configurations: config 3 (stages = 3), config 4 (5-bracket HDR
            fuse + develop); with N > 1 also ONE 100 MP frame over row bands and ONE 5-bracket HDR set with the
            brackets spread over the ranks, both through NVLink-shared memory (strong scaling), each checked
            element-wise against the single-GPU result;
  cpu_baseline (N=1, rank 0)  the unmodified reference (baseline/_ref/pySP, installed by oracle/install_ref.py) on one
            full 24 MP frame with all host threads; the oracle's NumPy/OpenCV port if the reference did not travel.
`--impl reference` times that CPU implementation alone, one full 24 MP frame per step, and prints the same JSON line.
Frames are sharded over ranks with no data-path collective (weak scaling).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpix/s AHD develop (24MP RGGB)"
H, W = 4000, 6000
STAGES = 1
FRAMES = 8                      # distinct frames per GPU: 8 x 48 MB of mosaic, larger than the 126 MB L2
MIN_TIMED_S = 2.0
ALGO_BYTES_PER_PX = {"ahd_select_kernel": 14.0, "median_stage_kernel": 24.0, "develop_chain": 14.0}
BIG_H, BIG_W = 8660, 11548      # BASELINE config 5: one 100 MP frame
# identical in both arms, so that the driver compares like with like
CONFIG = {"workload": "QualityDemosaic.Best AHD (postprocess_stages=1) + WB + cam->lin sRGB, synthetic 6000x4000 14-bit "
                      "RGGB, float32 out (BASELINE config 2)", "stages": STAGES, "frame": [W, H]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself when it travelled (baseline/_ref/pySP), else the oracle's NumPy/OpenCV port
def cpu_develop_fn(threads):
    """Returns (fn(raw_u16) -> seconds of one develop, kind, detail)."""
    os.environ["OMP_NUM_THREADS"] = str(threads)     # before the reference's OpenMP extension is loaded
    import cv2
    from pysp_b200 import synthetic as syn
    cv2.setNumThreads(threads)
    try:
        # torchrun exports OMP_NUM_THREADS=1 and NumPy's OpenBLAS read it when NumPy was imported: give the float64 matrix
        # products of the reference (np.dot, colorize/transform.py:52) every host thread back, as in a plain `python` run
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=threads)
    except Exception:
        pass
    try:
        from oracle import ref_harness as rh
        if not rh.available():
            raise RuntimeError("the reference install baseline/_ref/pySP is not in this snapshot "
                               "(oracle/install_ref.py runs in __graft_entry__.build() where /root/reference is mounted)")
        rh.load()
        rh.pin_numerics(False)                       # OpenCV default (optimised) mode: what a pySP user runs

        def ref_once(raw):
            t0 = time.perf_counter()
            rh.develop(raw, syn.BLACK, syn.WHITE, STAGES)
            return time.perf_counter() - t0

        return ref_once, "reference", ("unmodified bullbin/pySP imported from %s: bayer_normalize -> RawBayerData.demosaic("
                                       "Best, %d) -> to_lin_srgb(), OpenCV default mode, %d threads (cv2 + OpenMP)"
                                       % (os.path.relpath(rh.REFERENCE_ROOT, ROOT), STAGES, threads))
    except Exception as ex:
        why = "%s: %s" % (type(ex).__name__, str(ex)[:200])
    from oracle import ahd_spec as sp
    cv2.setUseOptimized(True)
    wb = syn.wb_multipliers()
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)

    def port_once(raw):
        t0 = time.perf_counter()
        sensor = sp.normalize(raw, syn.BLACK, syn.WHITE)
        cam = sp.ahd_demosaic(sensor, wb, m, STAGES, backend="cv2")
        sp.to_lin_srgb(cam, m, backend="cv2")
        return time.perf_counter() - t0

    return port_once, "port", "NumPy/OpenCV restatement of the reference (oracle cv2 backend); the reference itself was " \
                              "not usable: " + why


def run_reference(args, rank):
    if rank != 0:
        return
    from pysp_b200 import synthetic as syn
    threads = os.cpu_count() or 1
    once, kind, detail = cpu_develop_fn(threads)
    raw = syn.scene(H, W, 0)
    for _ in range(args.warmup):
        once(raw)
    ts = [once(raw) for _ in range(args.steps)]
    total = sum(ts)
    mpix = H * W * args.steps / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": CONFIG,
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": threads, "kind": kind, "detail": detail,
                             "sample": "one full 6000x4000 frame (seed 0) per step", "best_step_s": min(ts)},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--min-seconds", type=float, default=MIN_TIMED_S,
                    help="lower bound of the timed region (the 8-frame pass is looped to reach it); 0 for profiler runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from pysp_b200 import _capi, engine, parallel
    from pysp_b200 import synthetic as syn
    from pysp_b200.colour import cam_to_rgb_matrix
    from pysp_b200.pipeline import FramePipeline
    from pysp_b200.raw_hdr import fusion_constants
    from pysp_b200.wb_cct import CameraWhiteBalance
    import ctypes as C

    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        parallel.quiet_nccl_stdout()                                 # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    all_cpus = os.sched_getaffinity(0)
    numa = parallel.bind_to_gpu_numa_node(local_rank)   # pinned staging buffers of the e2e leg on the GPU's socket

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks_true(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def device_ms(fn, reps=1):
        """device time of `reps` calls of fn on the current stream (CUDA events), max over ranks"""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    cam_wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    wb = cam_wb.get_reciprocal_multipliers()
    m = cam_to_rgb_matrix(cam_wb.get_matrix())
    # frames of this rank's shard of the global batch (whole frames, round-robin: no collective on the data path)
    mine = parallel.frames_for_rank(FRAMES * world, rank, world)
    base = syn.scene_base(H, W)
    host = [torch.from_numpy(syn.scene(H, W, seed=i, base=base).view(np.int16)) for i in mine]
    del base
    frames = [h.to(dev) for h in host]
    outs = [torch.empty((H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    kw = dict(wb=wb, cam_to_srgb=m, black=syn.BLACK, white=syn.WHITE, out="lin")
    px_per_frame = H * W

    def one_pass(stages=STAGES):
        for i, f in enumerate(frames):
            engine.develop(f, out_tensor=outs[i & 1], stages=stages, **kw)

    # ---- headline: config 2, device-resident --------------------------------------------------------------------------
    one_pass()
    t_pass = device_ms(one_pass) * 1e-3                              # seconds per pass over the 8 frames, max over ranks
    loops = max(1, int(np.ceil(args.min_seconds / (args.steps * t_pass))))

    def step():
        for _ in range(loops):
            one_pass()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = engine.kernel_launches()
    lib.pysp_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join()
    tot = (C.c_double * 4)()
    cnt = (C.c_int64 * 4)()
    _capi.check(lib.pysp_timing_collect(tot, cnt))
    lib.pysp_timing_enable(0)
    launches = engine.kernel_launches() - launches0
    frames_per_step = len(frames) * loops
    value = world * frames_per_step * args.steps * px_per_frame / (ms * 1e-3) / 1e6

    # ---- end to end through the public batch API, host buffers ----------------------------------------------------------
    e2e, e2e_variants = None, {}
    if not args.no_e2e:
        pin_in = [h.pin_memory() for h in host]
        e2e_steps = max(20, args.steps)

        def e2e_leg(out_kind, steps, with_ceiling):
            pipe = FramePipeline(H, W, wb, m, stages=STAGES, black=syn.BLACK, white=syn.WHITE, out=out_kind, device=dev)
            pin_out = [pipe.pinned_output() for _ in range(len(host))]
            pipe.run(pin_in, pin_out)                    # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                pipe.run(pin_in, pin_out)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            res = {"value": world * len(host) * steps * px_per_frame / dt / 1e6, "unit": "Mpix/s",
                   "h2d_bytes_per_step": len(host) * pipe.h2d_bytes(), "d2h_bytes_per_step": len(host) * pipe.d2h_bytes(),
                   "steps": steps, "frames_per_step_per_gpu": len(host), "out": out_kind,
                   "d2h_GBps_per_gpu": len(host) * steps * pipe.d2h_bytes() / dt / 1e9}
            if with_ceiling:
                pipe.run_copies_only(pin_in, pin_out)
                barrier()
                t0 = time.perf_counter()
                for _ in range(steps):
                    pipe.run_copies_only(pin_in, pin_out)
                barrier()
                dc = max_over_ranks(time.perf_counter() - t0)
                res["ceiling_mpix_s"] = world * len(host) * steps * px_per_frame / dc / 1e6
                res["frac_of_ceiling"] = res["value"] / res["ceiling_mpix_s"]
                res["ceiling"] = "the same pinned buffers, byte counts, 3 streams and order, one cudaMemcpyAsync per copy, no kernels"
                pipe.run(pin_in, pin_out)                # the outputs are checked below
                barrier()
            return res, pin_out, pipe

        e2e, pin_out, pipe = e2e_leg("lin", e2e_steps, True)
        e2e.update({"api": "pysp_b200.pipeline.FramePipeline.run (pinned host in/out, 3 streams)", "numa_node": numa})
        # the pipeline produced the device-resident result
        ref0 = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        engine.develop(frames[0], out_tensor=ref0, stages=STAGES, **kw)
        torch.cuda.synchronize()
        assert torch.equal(pin_out[0].to(dev).view(torch.int32), ref0.view(torch.int32))
        del pipe, pin_out, ref0
        for kind in ("lin_f16", "srgb_u8"):
            r, po, pp = e2e_leg(kind, max(5, e2e_steps // 2), True)
            e2e_variants[kind] = r
            del po, pp
        del pin_in

    # ---- the other BASELINE configurations (device-resident) -----------------------------------------------------------
    workloads = {}
    if not args.no_workloads:
        # config 3: postprocess_stages = 3 on the same frames
        one_pass(3)
        reps = max(1, int(np.ceil(0.5 / max(device_ms(lambda: one_pass(3)) * 1e-3, 1e-6))))
        t = device_ms(lambda: one_pass(3), reps) * 1e-3
        workloads["config3_ahd_stages3_24MP"] = {"value": world * len(frames) * reps * px_per_frame / t / 1e6, "unit": "Mpix/s",
                                                 "frames": world * len(frames) * reps, "scaling": "weak"}
        # config 4: five 24 MP brackets fused in raw space + HDR develop, one set per GPU
        brackets, evs = syn.hdr_brackets(H, W, 5, 5)
        br = [engine.to_device(b, dev) for b in brackets]
        del brackets
        tev, offs, bias = fusion_constants(evs, wb)

        def hdr_set():
            fused, _ = engine.fuse_exposures(br, offs, bias, int(np.argmax(offs)), want_count=False)
            engine.develop(fused, wb, m, stages=STAGES, hdr=True, out_tensor=outs[0])

        hdr_set()
        reps = max(1, int(np.ceil(0.5 / max(device_ms(hdr_set) * 1e-3, 1e-6))))
        t = device_ms(hdr_set, reps) * 1e-3
        workloads["config4_hdr5_fuse_plus_ahd_24MP"] = {"value": world * reps * px_per_frame / t / 1e6, "unit": "Mpix/s (output pixels)",
                                                        "bracket_mpix_s": 5 * world * reps * px_per_frame / t / 1e6,
                                                        "sets": world * reps, "scaling": "weak"}
        if world > 1:
            workloads.update(multi_gpu_workloads(torch, dist, engine, parallel, syn, dev, rank, world, wb, m, br, evs, offs, bias,
                                                 device_ms, all_ranks_true, value / world))
        del br

    if rank == 0:
        peak, peak_src = peaks()
        names = ["ahd_select_kernel", "median_stage_kernel"]
        per = {names[k]: (tot[k] / cnt[k] if cnt[k] else 0.0) for k in range(2)}
        dom = max(per, key=lambda k: per[k])
        chain_ms = sum(per.values())
        ach = ALGO_BYTES_PER_PX[dom] * px_per_frame / (per[dom] * 1e-3) / 1e9 if per[dom] > 0 else 0.0
        traffic, issue, src = None, None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = prof.get(dom)
            src = "profiles/ncu_traffic.json, %s (constants from one ncu --set full capture, not measured in this run)" % prof.get("capture", "?")
            # the path is issue-bound, not HBM-bound: warp instructions per launch (ncu) over the
            # 148 SMs x 4 schedulers x 1 instruction/clock gives the floor the measured launch time is compared with
            clk = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
            issue = {k: {"warp_inst_per_launch": v, "thread_inst_per_px": 32.0 * v / px_per_frame,
                         "issue_floor_ms": 1e3 * v / (148 * 4 * clk),
                         "issue_frac": (1e3 * v / (148 * 4 * clk)) / per[k] if per.get(k) else None}
                     for k, v in prof.get("inst", {}).items()}
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX[dom] * px_per_frame,
                    "ms_per_launch": per, "launches": {names[k]: int(cnt[k]) for k in range(2)}, "issue": issue,
                    "issue_source": src,
                    "chain": {"bytes_per_px": 14.0, "ms_per_frame": chain_ms,
                              "achieved": 14.0 * px_per_frame / (chain_ms * 1e-3) / 1e9 if chain_ms else 0.0,
                              "frac": (14.0 * px_per_frame / (chain_ms * 1e-3) / 1e9 / peak) if chain_ms else 0.0}}
        cfg = dict(CONFIG)
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "run": {"frames_per_step_per_gpu": frames_per_step, "distinct_frames_per_gpu": len(frames), "loops_per_step": loops,
                        "timed_region_s": ms * 1e-3, "sharding": "whole frames, no collective",
                        "l2": "inputs (8 x 48 MB mosaics) and outputs (288 MB each) larger than the 126 MB L2"},
                "roofline": roofline, "clocks": sampler.summary(), "gpu_launches": int(launches)}
        if e2e is not None:
            line["e2e"] = e2e
            line["e2e_variants"] = e2e_variants
        if workloads:
            line["workloads"] = workloads
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)           # the CPU arm gets every host thread back
            threads = os.cpu_count() or 1
            once, kind, detail = cpu_develop_fn(threads)
            raw0 = syn.scene(H, W, 0)
            once(raw0[:1000, :1504])                    # loads the libraries, spins up the thread pools
            t = once(raw0)
            line["cpu_baseline"] = {"value": px_per_frame / t / 1e6, "unit": "Mpix/s", "cores": threads, "kind": kind,
                                    "detail": detail, "sample": "one full 6000x4000 frame (seed 0), after a warm-up on a 1504x1000 crop"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def multi_gpu_workloads(torch, dist, engine, parallel, syn, dev, rank, world, wb, m, br, evs, offs, bias, device_ms, all_ranks_true,
                        per_gpu_mpix_s):
    """N > 1: ONE frame / ONE bracket set spread over all ranks (strong scaling), through NVLink-shared memory
    (pysp_b200.parallel.SymmetricBand / SymmetricBrackets).  Every rank also develops the whole input alone and compares its
    band of the N-GPU result with the same rows of that single-GPU result, element by element."""
    res = {}
    stages = STAGES
    kw = dict(wb=wb, cam_to_srgb=m, stages=stages, black=syn.BLACK, white=syn.WHITE)
    # ---- one 100 MP frame over row bands; the raw halo rows are pulled out of the neighbours' HBM over NVLink ----
    try:
        frame = syn.scene(BIG_H, BIG_W, 1)
        b, e = parallel.band_rows(BIG_H, world, rank)
        my_rows = torch.from_numpy(frame[b:e].view(np.int16)).to(dev)
        sb, why = None, None
        try:                                              # NVLink-shared memory (a private torch API): agreed on by all ranks
            sb = parallel.SymmetricBand(BIG_H, BIG_W, torch.int16, stages)
            sb.band().copy_(my_rows)
        except Exception as ex:
            sb, why = None, repr(ex)[:200]
        symmetric = all_ranks_true(sb is not None)
        if symmetric:
            transport = "NVLink peer-to-peer copies of raw halo rows out of symmetric memory (no NCCL kernel)"

            def exchange():
                return sb.exchange()
        else:                                             # public-API fallback: NCCL send / recv of the halo rows
            sb = None
            transport = "NCCL send/recv of raw halo rows (symmetric memory unavailable: %s)" % why

            def exchange():
                held, hb = parallel.exchange_halo(my_rows, BIG_H, stages)
                padded = engine.alloc_rows(held.shape[0], BIG_W, torch.int16, dev)       # 16-byte aligned rows for TMA
                padded.copy_(held)
                return padded, hb

        def band_step():
            held, hb = exchange()
            return engine.develop(held, rows=(b, e), frame_height=BIG_H, in_row0=hb, **kw)

        out = band_step()
        whole_in = engine.to_device(frame.view(np.int16), dev, pad_pitch=True)
        whole = engine.develop(whole_in, **kw)                                            # this rank alone: the 1-GPU result
        same = all_ranks_true(torch.equal(out.view(torch.int32), whole[b:e].view(torch.int32)))
        reps = 5
        t1 = device_ms(lambda: engine.develop(whole_in, out_tensor=whole, **kw), reps) / reps
        del whole, out, whole_in
        t = device_ms(band_step, reps) / reps
        tx = device_ms(lambda: exchange(), reps) / reps
        mp = BIG_H * BIG_W / (t * 1e-3) / 1e6
        res["one_100MP_frame_row_bands"] = {
            "value": mp, "unit": "Mpix/s", "scaling": "strong", "ms_per_frame": t, "exchange_us": tx * 1e3,
            "bit_identical_to_1gpu": same, "compare": "element-wise, every rank its own band against its own whole-frame run",
            "transport": transport, "halo_rows": parallel.halo_rows(stages), "band_rows": e - b,
            "ms_per_frame_1gpu": t1, "strong_scaling_efficiency": t1 / (world * t),
            "efficiency_vs_frame_batch": mp / (per_gpu_mpix_s * world)}
        del sb, frame, my_rows
    except Exception as ex:
        res["one_100MP_frame_row_bands"] = {"error": repr(ex)[:300]}
    # ---- one 5-bracket HDR set, bracket k on rank k % world; the fuse kernel reads its operands from the owners' HBM ----
    try:
        nb = len(br)
        halo = parallel.halo_rows(stages)
        bb, be = parallel.band_rows(H, world, rank)
        sbr, why = None, None
        try:
            sbr = parallel.SymmetricBrackets(H, W, nb, halo)
            for k in range(nb):
                if sbr.owner(k) == rank:
                    sbr.slot(k).copy_(br[k])
        except Exception as ex:
            sbr, why = None, repr(ex)[:200]
        symmetric = all_ranks_true(sbr is not None)
        if symmetric:
            transport = ("fuse_kernel loads the brackets from their owners' HBM over NVLink (symmetric memory): exchange and "
                         "compute are one kernel")

            def hdr_step():
                sbr.ready()
                rows, hb = sbr.views()
                fused, _ = engine.fuse_exposures(rows, offs, bias, int(np.argmax(offs)), want_count=False)
                sbr.done()
                return engine.develop(fused, wb, m, stages=stages, hdr=True, rows=(bb, be), frame_height=H, in_row0=hb)

            def sync_only():
                sbr.ready()
                sbr.done()
        else:                                             # public-API fallback: every rank receives its band's rows of every bracket
            sbr = None
            transport = "NCCL send/recv of the band's rows of every bracket (symmetric memory unavailable: %s)" % why
            mine = {k: br[k] for k in range(nb) if k % world == rank}
            like = torch.empty((0, W), dtype=torch.float32, device=dev)

            def gather_rows():
                return parallel.exchange_brackets_by_rows(mine, nb, H, halo, like=like)

            def hdr_step():
                rows, hb = gather_rows()
                fused, _ = engine.fuse_exposures(rows, offs, bias, int(np.argmax(offs)), want_count=False)
                return engine.develop(fused, wb, m, stages=stages, hdr=True, rows=(bb, be), frame_height=H, in_row0=hb)

            sync_only = gather_rows

        out = hdr_step()
        # the single-GPU result of the same set (every rank holds a copy of all five brackets in `br`)
        f1, _ = engine.fuse_exposures(br, offs, bias, int(np.argmax(offs)), want_count=False)
        whole = engine.develop(f1, wb, m, stages=stages, hdr=True)
        same = all_ranks_true(torch.equal(out.view(torch.int32), whole[bb:be].view(torch.int32)))
        del whole, f1, out
        reps = 10
        t = device_ms(hdr_step, reps) / reps
        tb = device_ms(sync_only, reps) / reps
        res["one_hdr5_set_brackets_sharded"] = {
            "value": H * W / (t * 1e-3) / 1e6, "unit": "Mpix/s (output pixels)", "scaling": "strong", "ms_per_set": t,
            "bit_identical_to_1gpu": same, "compare": "element-wise, every rank its own band against its own single-GPU run",
            "transport": transport, "exchange_us": 0.0 if symmetric else tb * 1e3,
            "device_barriers_us": tb * 1e3 if symmetric else None}
        del sbr
    except Exception as ex:
        res["one_hdr5_set_brackets_sharded"] = {"error": repr(ex)[:300]}
    return res


if __name__ == "__main__":
    main()
