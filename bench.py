#!/usr/bin/env python
"""Benchmark of the develop hot path (BASELINE.json metric: Mpix/s of AHD develop on 24 MP RGGB).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" develops one batch of FRAMES distinct synthetic 6000x4000 14-bit RGGB frames per GPU
(BASELINE config 2: QualityDemosaic.Best, postprocess_stages=1, WB + camera->linear sRGB, float32 out).
  value     whole-job Mpix/s with the mosaics already resident in HBM (CUDA events, max over ranks);
  e2e       the same metric through the public batch API (pysp_b200.pipeline.FramePipeline) with HOST
            buffers: pinned H2D of every mosaic and D2H of every result inside the timed region;
  roofline  dominant kernel: algorithmic bytes per launch / its mean device time (CUDA events recorded by
            the library around each launch on the launching stream) against the measured HBM copy peak;
  cpu_baseline (N=1, rank 0)  the oracle port with the reference's own library calls (NumPy + OpenCV, all
            host threads) on a bounded crop of the same frame.
`--impl reference` times that CPU port alone, one bounded sample per step, and prints the same JSON line.
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
FRAMES = 8                      # per GPU per step: 8 x 48 MB of mosaic, larger than the 126 MB L2
ALGO_BYTES_PER_PX = {"ahd_select_kernel": 14.0, "median_stage_kernel": 24.0, "develop_chain": 14.0}
CPU_SAMPLE = (2000, 3000)       # bounded crop for the CPU baseline (6 MP)


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


def cpu_port(sample_hw, threads):
    """One develop of a bounded crop with the oracle's cv2 backend (the reference's own library calls);
    the reference's compiled count map (oracle/_ref) is used when it travelled with the repo."""
    import cv2
    from oracle import ahd_spec as sp
    from pysp_b200 import synthetic as syn
    cv2.setNumThreads(threads)
    cv2.setUseOptimized(True)
    count_fn, kind = None, "port: NumPy/OpenCV restatement of the reference (oracle cv2 backend)"
    so_dir = os.path.join(ROOT, "oracle", "_ref")
    try:
        import importlib.machinery
        import importlib.util
        so = [f for f in os.listdir(so_dir) if f.startswith("ahd_homogeneity_cython") and f.endswith(".so")]
        if so:
            path = os.path.join(so_dir, so[0])
            loader = importlib.machinery.ExtensionFileLoader("ahd_homogeneity_cython", path)
            spec = importlib.util.spec_from_file_location("ahd_homogeneity_cython", path, loader=loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            count_fn = lambda lab, vertical: mod.build_map(np.ascontiguousarray(lab), 1, 3, bool(vertical))  # noqa: E731
            kind = "port: NumPy/OpenCV restatement of the reference + the reference's own compiled count map (oracle/_ref)"
    except Exception:
        count_fn = None
    raw = syn.scene(sample_hw[0], sample_hw[1], 0)
    wb = syn.wb_multipliers()
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)

    def once():
        t0 = time.perf_counter()
        sensor = sp.normalize(raw, syn.BLACK, syn.WHITE)
        cam = sp.ahd_demosaic(sensor, wb, m, STAGES, backend="cv2", count_fn=count_fn)
        sp.to_lin_srgb(cam, m, backend="cv2")
        return time.perf_counter() - t0

    return once, kind


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)     # torchrun pins it to 1; the CPU arm uses every host thread
    once, kind = cpu_port(CPU_SAMPLE, threads)
    for _ in range(args.warmup):
        once()
    ts = [once() for _ in range(args.steps)]
    total = sum(ts)
    mpix = CPU_SAMPLE[0] * CPU_SAMPLE[1] * args.steps / total / 1e6
    sample = "one %dx%d crop of the 24 MP frame per step" % (CPU_SAMPLE[1], CPU_SAMPLE[0])
    line = {"impl": "reference", "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "QualityDemosaic.Best AHD (postprocess_stages=1) + WB + cam->lin sRGB, synthetic "
                                   "6000x4000 14-bit RGGB; CPU arm on a bounded crop", "stages": STAGES},
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": threads, "kind": "port", "detail": kind, "sample": sample},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
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

    cam_wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    wb = cam_wb.get_reciprocal_multipliers()
    m = cam_to_rgb_matrix(cam_wb.get_matrix())
    # frames of this rank's shard of the global batch (whole frames, round-robin: no collective on the data path)
    mine = parallel.frames_for_rank(FRAMES * world, rank, world)
    base = syn.scene_base(H, W)
    host = [torch.from_numpy(syn.scene(H, W, seed=i, base=base).view(np.int16)) for i in mine]
    frames = [h.to(dev) for h in host]
    outs = [torch.empty((H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    kw = dict(wb=wb, cam_to_srgb=m, stages=STAGES, black=syn.BLACK, white=syn.WHITE, out="lin")

    def step():
        for i, f in enumerate(frames):
            engine.develop(f, out_tensor=outs[i & 1], **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    tot = (C.c_double * 4)()
    cnt = (C.c_int64 * 4)()
    _capi.check(lib.pysp_timing_collect(tot, cnt))
    lib.pysp_timing_enable(0)
    launches = engine.kernel_launches() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    px_per_frame = H * W
    value = world * len(frames) * args.steps * px_per_frame / (ms * 1e-3) / 1e6

    # ---- end to end through the public batch API, host buffers ----
    e2e = None
    if not args.no_e2e:
        pipe = FramePipeline(H, W, wb, m, stages=STAGES, black=syn.BLACK, white=syn.WHITE, out="lin", device=dev)
        pin_in = [h.pin_memory() for h in host]
        pin_out = [pipe.pinned_output() for _ in range(len(host))]
        pipe.run(pin_in, pin_out)                    # warm-up
        barrier()
        e2e_steps = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            pipe.run(pin_in, pin_out)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * len(host) * e2e_steps * px_per_frame / float(t.item()) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": len(host) * pipe.h2d_bytes(), "d2h_bytes_per_step": len(host) * pipe.d2h_bytes(),
               "steps": e2e_steps, "api": "pysp_b200.pipeline.FramePipeline.run (pinned host in/out, 3 streams)", "numa_node": numa}
        # light check that the pipeline produced the device-resident result
        ref0 = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        engine.develop(frames[0], out_tensor=ref0, **kw)
        torch.cuda.synchronize()
        assert torch.equal(pin_out[0].to(dev).view(torch.int32), ref0.view(torch.int32))
        del pipe, pin_in, pin_out, ref0

    if rank == 0:
        peak, peak_src = peaks()
        names = ["ahd_select_kernel", "median_stage_kernel"]
        per = {names[k]: (tot[k] / cnt[k] if cnt[k] else 0.0) for k in range(2)}
        dom = max(per, key=lambda k: per[k])
        chain_ms = sum(per.values())
        ach = ALGO_BYTES_PER_PX[dom] * px_per_frame / (per[dom] * 1e-3) / 1e9 if per[dom] > 0 else 0.0
        traffic, issue = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = prof.get(dom)
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
                    "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX[dom] * px_per_frame,
                    "ms_per_launch": per, "launches": {names[k]: int(cnt[k]) for k in range(2)}, "issue": issue,
                    "chain": {"bytes_per_px": 14.0, "ms_per_frame": chain_ms,
                              "achieved": 14.0 * px_per_frame / (chain_ms * 1e-3) / 1e9 if chain_ms else 0.0,
                              "frac": (14.0 * px_per_frame / (chain_ms * 1e-3) / 1e9 / peak) if chain_ms else 0.0}}
        line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "QualityDemosaic.Best AHD (postprocess_stages=1) + WB + cam->lin sRGB, synthetic "
                                       "6000x4000 14-bit RGGB, float32 out (BASELINE config 2)",
                           "frames_per_step_per_gpu": len(frames), "stages": STAGES, "sharding": "whole frames, no collective",
                           "l2": "inputs (8 x 48 MB mosaics) and outputs (288 MB each) larger than the 126 MB L2"},
                "roofline": roofline, "clocks": sampler.summary(), "gpu_launches": int(launches)}
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            os.sched_setaffinity(0, all_cpus)           # the CPU arm gets every host thread back
            threads = os.cpu_count() or 1
            os.environ["OMP_NUM_THREADS"] = str(threads)
            once, kind = cpu_port(CPU_SAMPLE, threads)
            once()
            best = min(once() for _ in range(2))
            line["cpu_baseline"] = {"value": CPU_SAMPLE[0] * CPU_SAMPLE[1] / best / 1e6, "unit": "Mpix/s", "cores": threads,
                                    "kind": "port", "detail": kind,
                                    "sample": "%dx%d crop of frame 0, best of 2 after 1 warm-up" % (
                                        CPU_SAMPLE[1], CPU_SAMPLE[0])}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
