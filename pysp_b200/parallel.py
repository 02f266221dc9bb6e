"""Multi-GPU partitioning of the develop path: one process per GPU, torch.distributed for plumbing.

The reference has no distributed code; the path shards naturally:
  * batches of frames / HDR sets  -> whole frames round-robin over ranks, NO data-path collective;
  * one very large frame          -> contiguous, even-aligned row bands; each rank needs 6 + 4*stages
                                     mosaic rows of its neighbours: one halo exchange of RAW rows between
                                     adjacent ranks (NCCL send/recv = NVLink P2P), recompute in the halo,
                                     never exchange intermediates; the result is bit-identical to 1 GPU;
  * HDR brackets spread over ranks -> one exchange step (every rank pulls its band's rows of every
                                     bracket from the owners) so that each rank accumulates the brackets
                                     in list order, like raw_hdr.py:135-139, bit for bit.
The exchange helpers work on CUDA tensors with the NCCL backend and on CPU tensors with gloo (tests).
"""
import os

import torch
import torch.distributed as dist


def quiet_nccl_stdout():
    """NCCL_DEBUG=VERSION makes NCCL print its version banner on stdout; tools that print one JSON line keep stdout clean."""
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: run this process (and first-touch its pinned staging buffers) on the CPU socket the GPU's PCIe
    root hangs off, so that host<->device copies do not cross the socket interconnect.  Best effort; returns the NUMA node
    or None when the topology cannot be read."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def _symmetric_memory_allowed():
    """`torch.distributed._symmetric_memory` is a private torch API; PYSP_NO_SYMMETRIC_MEMORY=1 disables its use (callers
    then take the NCCL transports `exchange_halo` / `exchange_brackets_by_rows`, which give the same bits)."""
    if os.environ.get("PYSP_NO_SYMMETRIC_MEMORY"):
        raise RuntimeError("symmetric memory disabled by PYSP_NO_SYMMETRIC_MEMORY")


def frames_for_rank(n_frames, rank, world):
    """Whole-frame sharding: frame i goes to rank i % world."""
    return list(range(rank, n_frames, world))


def band_rows(height, world, rank):
    """Even-aligned contiguous row band [begin, end) of `rank`; bands tile [0, height)."""
    quads = height // 2
    base, extra = divmod(quads, world)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return 2 * begin, 2 * end


def halo_rows(stages):
    return 6 + 4 * max(int(stages), 0)


def band_with_halo(height, world, rank, stages):
    """(band_begin, band_end, held_begin, held_end): rows a rank must hold to develop its band."""
    b, e = band_rows(height, world, rank)
    h = halo_rows(stages)
    return b, e, max(0, b - h), min(height, e + h)


def _bytes(t):
    """Raw rows travel as bytes: NCCL has no 16-bit integer type, and the payload is opaque anyway."""
    return t.view(torch.uint8) if t.dtype in (torch.int16, torch.uint16) else t


def exchange_halo(band, height, stages, group=None):
    """band: [rows, W] tensor holding exactly this rank's band rows.  Returns ([held_rows, W] tensor,
    held_begin): the band extended by the neighbours' raw rows.  Bands shorter than the halo pull from
    ranks further away, so every rank sends each peer exactly the rows that peer needs."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b, e, hb, he = band_with_halo(height, world, rank, stages)
    assert band.shape[0] == e - b, (band.shape, b, e)
    held = torch.empty((he - hb,) + tuple(band.shape[1:]), dtype=band.dtype, device=band.device)
    held[b - hb:e - hb].copy_(band)
    ops, keep = [], []
    for peer in range(world):
        if peer == rank:
            continue
        pb, pe, phb, phe = band_with_halo(height, world, peer, stages)
        # rows of mine that the peer needs
        s0, s1 = max(b, phb), min(e, phe)
        if s0 < s1:
            t = _bytes(band[s0 - b:s1 - b].contiguous())
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, peer, group))
        # rows of the peer that I need
        r0, r1 = max(pb, hb), min(pe, he)
        if r0 < r1:
            ops.append(dist.P2POp(dist.irecv, _bytes(held[r0 - hb:r1 - hb]), peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return held, hb


class SymmetricBand:
    """Row band of one large frame in NVLink-shared ("symmetric") memory: every rank allocates the same buffer, maps its
    peers' buffers (torch.distributed._symmetric_memory: CUDA IPC over NVLink / NVSwitch) and *pulls* the raw halo rows
    it needs straight out of its neighbours' HBM with peer-to-peer copies -- no NCCL kernel, no staging copy of the band.

        sb = SymmetricBand(height, width, torch.int16, stages)      # collective: all ranks of the group
        sb.band().copy_(my_rows)                                     # or decode / upload straight into sb.band()
        held, held_begin = sb.exchange()                             # stream-ordered; held = band +- halo rows
        out = engine.develop(held, ..., rows=sb.rows, frame_height=height, in_row0=held_begin)

    Results are bit-identical to the NCCL path (`exchange_halo`) and to one GPU.  CUDA only; the CPU tests cover the
    same row arithmetic through `exchange_halo` on gloo."""

    def __init__(self, height, width, dtype, stages, group=None, device=None):
        _symmetric_memory_allowed()
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.height, self.width, self.stages, self.dtype = int(height), int(width), int(stages), dtype
        self.ranges = [band_with_halo(self.height, self.world, r, self.stages) for r in range(self.world)]
        b, e, hb, he = self.ranges[self.rank]
        self.rows, self.held_begin, self.held_rows = (b, e), hb, he - hb
        self.max_rows = max(r[3] - r[2] for r in self.ranges)
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # row pitch padded to 128 bytes: the develop kernels move 16-byte aligned rows with TMA (11548 x 2 B is not aligned)
        from .engine import padded_pitch_elems
        self.pitch = padded_pitch_elems(self.width, torch.empty((), dtype=dtype).element_size())
        self.store = symm.empty((self.max_rows, self.pitch), dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.store, group)
        self.buf = self.store[:, :self.width]
        self.peers = {}
        for peer in range(self.world):
            if peer != self.rank:
                pb, pe, phb, phe = self.ranges[peer]
                r0, r1 = max(pb, hb), min(pe, he)          # rows of the peer's band that I hold as halo
                if r0 < r1:
                    view = self.handle.get_buffer(peer, (self.max_rows, self.pitch), dtype, 0)
                    # whole padded rows travel: one contiguous peer-to-peer copy per neighbour
                    self.peers[peer] = (view[r0 - phb:r1 - phb], self.store[r0 - hb:r1 - hb])

    def band(self):
        """[band_rows, W] view of the shared buffer where this rank's own rows live."""
        b, e = self.rows
        return self.buf[b - self.held_begin:e - self.held_begin]

    def exchange(self):
        """Pull the neighbours' halo rows (peer-to-peer copies on the current stream, fenced by device-side barriers)."""
        self.handle.barrier(channel=0)                      # every rank's band is in place
        for src, dst in self.peers.values():
            dst.copy_(src, non_blocking=True)
        self.handle.barrier(channel=1)                      # nobody overwrites a band that is still being read
        return self.buf[:self.held_rows], self.held_begin


class SymmetricBrackets:
    """HDR brackets of one scene spread over the ranks (bracket k lives on rank k % world) in NVLink-shared memory.
    `views(rank_rows)` returns, for every bracket in list order, a tensor over the OWNER's HBM restricted to this rank's
    rows: handing those to `engine.fuse_exposures` makes the fuse kernel read its operands over NVLink / NVSwitch while
    it accumulates them in list order -- exchange and compute are one kernel, nothing is staged, and the float32 sums
    are those of raw_hdr.py:135-139 bit for bit.

        sbr = SymmetricBrackets(height, width, n_brackets, halo_rows(stages))   # collective
        sbr.slot(k).copy_(bracket_k)          # on the owner of bracket k
        rows, held_begin = sbr.views()        # after sbr.ready()
    """

    def __init__(self, height, width, n_brackets, halo, group=None, device=None, dtype=torch.float32):
        _symmetric_memory_allowed()
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.height, self.width, self.n = int(height), int(width), int(n_brackets)
        self.slots = (self.n + self.world - 1) // self.world
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((self.slots, self.height, self.width), dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.buf, group)
        b, e = band_rows(self.height, self.world, self.rank)
        self.rows = (b, e)
        self.held = (max(0, b - halo), min(self.height, e + halo))
        self._peer = {r: (self.buf if r == self.rank else
                          self.handle.get_buffer(r, (self.slots, self.height, self.width), dtype, 0)) for r in range(self.world)}

    def owner(self, k):
        return k % self.world

    def slot(self, k):
        """[H, W] view of bracket k on its owner (call on rank owner(k))."""
        assert self.owner(k) == self.rank
        return self.buf[k // self.world]

    def ready(self):
        """Device-side barrier on the current stream: every owner's brackets are in place."""
        self.handle.barrier(channel=0)

    def done(self):
        """Device-side barrier: every rank has finished reading (call before brackets are overwritten)."""
        self.handle.barrier(channel=1)

    def views(self):
        hb, he = self.held
        return [self._peer[self.owner(k)][k // self.world, hb:he] for k in range(self.n)], hb


def exchange_brackets_by_rows(my_brackets, n_brackets, height, halo, group=None, like=None):
    """HDR brackets live on rank (index % world) as whole [H, W] float32 mosaics (`my_brackets` maps
    bracket index -> tensor).  Returns the list of all n brackets restricted to this rank's rows
    [held_begin, held_end) = band +- halo, in list order, plus held_begin.  A rank that owns no
    bracket (world > n) passes `like`, an empty [0, W] tensor of the right dtype/device."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    any_t = next(iter(my_brackets.values())) if my_brackets else like
    width = any_t.shape[1]

    def held_range(r):
        b, e = band_rows(height, world, r)
        return max(0, b - halo), min(height, e + halo)

    hb, he = held_range(rank)
    out = [None] * n_brackets
    ops, keep = [], []
    for k in range(n_brackets):
        owner = k % world
        if owner == rank:
            out[k] = my_brackets[k][hb:he].contiguous()
            for peer in range(world):
                if peer != rank:
                    pb, pe = held_range(peer)
                    t = my_brackets[k][pb:pe].contiguous()
                    keep.append(t)
                    ops.append(dist.P2POp(dist.isend, t, peer, group))
        else:
            out[k] = torch.empty((he - hb, width), dtype=any_t.dtype, device=any_t.device)
            ops.append(dist.P2POp(dist.irecv, out[k], owner, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out, hb


def develop_band(band_mosaic, height, stages, develop_fn, group=None):
    """Single large frame over row bands: exchange raw halo rows, then develop this rank's band.
    develop_fn(held, in_row0, rows) -> the band's output (engine.develop on the GPU)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    held, hb = exchange_halo(band_mosaic, height, stages, group)
    b, e = band_rows(height, world, rank)
    return develop_fn(held, hb, (b, e))
