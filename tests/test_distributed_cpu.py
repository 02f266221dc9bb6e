"""N>1 host logic on CPU: two gloo ranks exchange raw halo rows / HDR bracket rows, develop their band
through the host emulation of the kernels, and must reproduce the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

H, W, STAGES = 48, 40, 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pysp_b200 import parallel
        from pysp_b200 import synthetic as syn
        import test_tile_logic as tl
        import ctypes as C
        from pysp_b200 import _capi
        lib = C.CDLL(os.path.join(ROOT, "tests", "host_emu", "libpysp_emu.so"))
        lib.emu_develop.argtypes = [C.POINTER(_capi.DevelopArgs), C.c_int, C.c_int]
        lib.emu_last_error.restype = C.c_char_p
        raw = syn.scene(H, W, 21)
        # --- single frame over row bands: halo exchange of the raw mosaic ---
        b, e = parallel.band_rows(H, world, rank)
        mine = torch.from_numpy(raw[b:e].view(np.int16).copy())

        def dev(held, hb, rows):
            full = np.zeros((H, W), dtype=np.uint16)
            full[hb:hb + held.shape[0]] = held.numpy().view(np.uint16)
            return tl.emu_develop(lib, full, STAGES, band=rows, held=(hb, held.shape[0]))

        band_out = parallel.develop_band(mine, H, STAGES, dev)
        np.save(os.path.join(out_dir, "band%d.npy" % rank), band_out)
        # --- HDR brackets owned round-robin: exchange by rows, fuse in list order ---
        from oracle import ahd_spec as sp
        base = (raw.astype(np.float32) - 512.0) / 16383.0
        brackets = [np.clip(base * np.float32(2.0 ** (-k)), 0, 1).astype(np.float32) for k in range(-1, 2)]
        evs = [9.0, 10.0, 11.0]
        mineb = {k: torch.from_numpy(brackets[k]) for k in range(3) if k % world == rank}
        rows, hb = parallel.exchange_brackets_by_rows(mineb, 3, H, 6, like=torch.empty((0, W)))
        fused, cnt, lim, tev = sp.fuse_exposures([r.numpy() for r in rows], evs, syn.wb_multipliers())
        np.save(os.path.join(out_dir, "fuse%d.npy" % rank), fused)
        np.save(os.path.join(out_dir, "fuse_hb%d.npy" % rank), np.array([hb]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_bands_and_bracket_exchange(tmp_path):
    import test_tile_logic as tl
    if not os.path.exists(os.path.join(ROOT, "tests", "host_emu", "libpysp_emu.so")):
        import subprocess
        subprocess.check_call([os.path.join(ROOT, "tests", "host_emu", "build.sh")])
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import ctypes as C
    from pysp_b200 import _capi, parallel
    from pysp_b200 import synthetic as syn
    from oracle import ahd_spec as sp
    lib = C.CDLL(os.path.join(ROOT, "tests", "host_emu", "libpysp_emu.so"))
    lib.emu_develop.argtypes = [C.POINTER(_capi.DevelopArgs), C.c_int, C.c_int]
    lib.emu_last_error.restype = C.c_char_p
    raw = syn.scene(H, W, 21)
    whole = tl.emu_develop(lib, raw, STAGES)
    got = np.concatenate([np.load(os.path.join(str(tmp_path), "band%d.npy" % r)) for r in range(world)])
    assert np.array_equal(got.view(np.uint32), whole.view(np.uint32))
    lin, _ = sp.develop(raw, syn.BLACK, syn.WHITE, syn.wb_multipliers(), syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, STAGES)
    assert np.array_equal(got.view(np.uint32), lin.view(np.uint32))
    base = (raw.astype(np.float32) - 512.0) / 16383.0
    brackets = [np.clip(base * np.float32(2.0 ** (-k)), 0, 1).astype(np.float32) for k in range(-1, 2)]
    fused, _, _, _ = sp.fuse_exposures(brackets, [9.0, 10.0, 11.0], syn.wb_multipliers())
    for r in range(world):
        part = np.load(os.path.join(str(tmp_path), "fuse%d.npy" % r))
        hb = int(np.load(os.path.join(str(tmp_path), "fuse_hb%d.npy" % r))[0])
        assert np.array_equal(part.view(np.uint32), fused[hb:hb + part.shape[0]].view(np.uint32))
