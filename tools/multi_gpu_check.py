"""N-GPU check (torchrun, NCCL): (1) one large frame over row bands with a halo exchange of raw rows between
neighbours, (2) HDR brackets owned round-robin, exchanged by rows and fused in list order; both must reproduce the
single-GPU result bit for bit.  Prints one JSON line from rank 0.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pysp_b200 import engine, parallel, synthetic as syn  # noqa: E402
from pysp_b200.colour import cam_to_rgb_matrix  # noqa: E402
from pysp_b200.raw_hdr import fusion_constants  # noqa: E402
from pysp_b200.wb_cct import CameraWhiteBalance  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    parallel.quiet_nccl_stdout()
    parallel.bind_to_gpu_numa_node(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    wbc = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    wb, m = wbc.get_reciprocal_multipliers(), cam_to_rgb_matrix(wbc.get_matrix())
    H, W, stages = 8660, 11548, 1                      # BASELINE config 5: one 100 MP frame
    small = syn.scene(H // 4 + 2, W // 4 + 2, 3)
    frame = np.tile(small, (4, 4))[:H, :W].copy()
    kw = dict(wb=wb, cam_to_srgb=m, stages=stages, black=syn.BLACK, white=syn.WHITE)
    # --- row bands: every rank starts with ONLY its band of the mosaic on its GPU ---
    b, e = parallel.band_rows(H, world, rank)
    band = torch.from_numpy(frame[b:e].view(np.int16)).to(dev)
    def device_ms(fn, reps=3):
        """max over ranks of the device time of fn (CUDA events on the current stream), best of `reps` after a warm-up"""
        fn()
        best = None
        for _ in range(reps):
            torch.cuda.synchronize(); dist.barrier()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b_.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b_)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return r, best

    out, t_band = device_ms(lambda: parallel.develop_band(band, H, stages, lambda held, hb, rows: engine.develop(
        held, rows=rows, frame_height=H, in_row0=hb, **kw)))
    _, t_exch = device_ms(lambda: parallel.exchange_halo(band, H, stages))
    t_band, t_exch = t_band * 1e-3, t_exch * 1e-3
    # the same band in NVLink-shared memory: halo rows pulled peer-to-peer, no NCCL kernel
    symm = {"available": False}
    try:
        sb = parallel.SymmetricBand(H, W, torch.int16, stages)
        sb.band().copy_(band)

        def symm_step():
            held, hb = sb.exchange()
            return engine.develop(held, rows=sb.rows, frame_height=H, in_row0=hb, **kw)

        out_s, t_s = device_ms(symm_step)
        _, t_sx = device_ms(lambda: sb.exchange())
        same = torch.tensor([int(torch.equal(out_s.view(torch.int32), out.view(torch.int32)))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        symm = {"available": True, "bit_identical_to_nccl_path": bool(same.item()), "bands_100MP_s": t_s * 1e-3,
                "bands_100MP_Mpix_s": H * W / (t_s * 1e-3) / 1e6, "halo_exchange_s": t_sx * 1e-3}
    except Exception as ex:                               # symmetric memory needs P2P-capable GPUs and driver support
        symm = {"available": False, "error": repr(ex)[:200]}
    # reference: EVERY rank develops the whole frame alone and compares its own band element by element
    whole = engine.develop(torch.from_numpy(frame.view(np.int16)).to(dev), **kw)
    same = torch.tensor([int(torch.equal(whole[b:e].view(torch.int32), out.view(torch.int32)))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ok_bands = bool(same.item())
    del whole
    # --- HDR brackets spread over ranks (config 4) ---
    Hh, Wh, nb = 4000, 6000, 5
    base = (np.tile(syn.scene(Hh // 4, Wh // 4, 5), (4, 4)).astype(np.float32) - 512.0) / 16383.0
    evs = [8.0, 9.0, 10.0, 11.0, 12.0]
    mine = {}
    for k in range(nb):
        if k % world == rank:
            mine[k] = torch.from_numpy(np.clip(base * np.float32(2.0 ** (2 - k)), 0, 1).astype(np.float32)).to(dev)
    halo = parallel.halo_rows(stages)
    tev, offs, bias = fusion_constants(evs, wb)
    bb, be = parallel.band_rows(Hh, world, rank)

    def hdr_step():
        rows, hb = parallel.exchange_brackets_by_rows(mine, nb, Hh, halo, like=torch.empty((0, Wh), dtype=torch.float32, device=dev))
        fused, _ = engine.fuse_exposures(rows, offs, bias, int(np.argmax(offs)), want_count=False)
        return engine.develop(fused, wb, m, stages=stages, hdr=True, rows=(bb, be), frame_height=Hh, in_row0=hb)

    hdr_out, t_hdr = device_ms(hdr_step)
    t_hdr *= 1e-3
    # the same brackets in NVLink-shared memory: the fuse kernel reads its operands straight from the owners' HBM
    symm_hdr = {"available": False}
    try:
        sbr = parallel.SymmetricBrackets(Hh, Wh, nb, halo)
        for k, t in mine.items():
            sbr.slot(k).copy_(t)

        def hdr_symm():
            sbr.ready()
            rows, hb = sbr.views()
            fused, _ = engine.fuse_exposures(rows, offs, bias, int(np.argmax(offs)), want_count=False)
            sbr.done()
            return engine.develop(fused, wb, m, stages=stages, hdr=True, rows=(bb, be), frame_height=Hh, in_row0=hb)

        out_s, t_s = device_ms(hdr_symm)
        same = torch.tensor([int(torch.equal(out_s.view(torch.int32), hdr_out.view(torch.int32)))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        symm_hdr = {"available": True, "bit_identical_to_nccl_path": bool(same.item()), "hdr5_24MP_s": t_s * 1e-3}
    except Exception as ex:
        symm_hdr = {"available": False, "error": repr(ex)[:200]}
    # every rank fuses and develops the whole set alone and compares its own band element by element
    allb = [torch.from_numpy(np.clip(base * np.float32(2.0 ** (2 - k)), 0, 1).astype(np.float32)).to(dev) for k in range(nb)]
    f1, _ = engine.fuse_exposures(allb, offs, bias, int(np.argmax(offs)), want_count=False)
    whole = engine.develop(f1, wb, m, stages=stages, hdr=True)
    same = torch.tensor([int(torch.equal(whole[bb:be].view(torch.int32), hdr_out.view(torch.int32)))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    ok_hdr = bool(same.item())
    del whole, f1, allb
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bands_100MP_bit_identical": bool(ok_bands), "bands_100MP_s": t_band,
                          "bands_100MP_Mpix_s": H * W / t_band / 1e6, "halo_exchange_s": t_exch,
                          "timing": "CUDA events, max over ranks, best of 3", "symmetric_memory": symm, "hdr5_24MP_bit_identical": bool(ok_hdr),
                          "hdr5_24MP_s": t_hdr, "hdr_symmetric_memory": symm_hdr}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
