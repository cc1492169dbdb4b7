"""Component sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Every component of a SkyModel is independent (no reduction, no neighbour access), so the hot path
shards trivially: rank r owns a contiguous block of components (for HEALPix maps a contiguous pixel
range whose centres are generated locally) and runs ``update_positions`` / ``coherency_calc`` on it with
no data-path collective.  The per-time context, R_avg, the frequency axis and the spline operator are
replicated (a few hundred bytes to ~10 KB).

Only two exchanges exist, both optional and outside the compute path:
  * :func:`exchange_counts` -- all-gather of the per-rank above-horizon counts (world_size int64) giving
    every rank the column offset of its block in the global compacted array, i.e. what the reference's
    single boolean index ``frame_coherency[..., above_horizon]`` (skymodel.py:2685) implies;
  * :func:`gather_columns` -- assemble the global (2, 2, Nfreqs, sum N_above) array on every rank over
    NCCL/NVLink.  Its volume is the full output per receiving GPU (tens of GB), ~50x the compute time,
    so it is never part of the timed hot path.

  * :class:`PeerGather` -- the B200-native form of the same assembly: the gathered array lives in symmetric
    (peer-mapped) memory and every rank's coherency kernel *writes its column range directly into the
    destination GPU's HBM* through NVLink (``ld_out`` / ``col0`` of ``rsk_coherency``), so the transfer is the
    kernel's own stores, tile by tile, and no data collective runs at all.

Works with the ``nccl`` backend on CUDA tensors and with ``gloo`` on CPU tensors (used by the CPU tests).
"""

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous block [lo, hi) of rank ``rank``: blocks of ceil(n / world_size) components."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("need 0 <= rank < world_size")
    per = -(-int(n) // world_size)
    lo = min(int(n), rank * per)
    return lo, min(int(n), lo + per)


def shard_kwargs(kwargs, world_size, rank):
    """Slice every per-component array of a ``SkyModel(**kwargs)`` call to this rank's block.

    Per-component arguments: lon/lat/ra/dec, stokes (last axis), spectral_index, reference_frequency,
    hpx_inds, name.  For an implicit full HEALPix map (nside given, no hpx_inds) ``hpx_first`` is set.
    """
    out = dict(kwargs)
    n = np.shape(kwargs["stokes"])[-1]
    lo, hi = shard_bounds(n, world_size, rank)
    for key in ("lon", "lat", "ra", "dec", "spectral_index", "reference_frequency", "hpx_inds", "name"):
        if out.get(key) is not None:
            out[key] = out[key][lo:hi]
    out["stokes"] = kwargs["stokes"][..., lo:hi]
    if out.get("nside") is not None and out.get("hpx_inds") is None:
        out["hpx_first"] = int(kwargs.get("hpx_first", 0)) + lo
    return out


def shard_sky(sky, world_size, rank):
    """This rank's contiguous block of an existing :class:`SkyModel` (e.g. one every rank loaded from the same SkyH5 or
    GLEAM file): ``sky.select(component_inds=[lo, hi))`` on a copy, so every per-component array is cut on the GPU by
    the row-gather kernel.  The blocks of all ranks, concatenated in rank order (``SkyModel.concat``), give the
    original catalogue back.  An empty block (more ranks than components) raises, as ``select`` does upstream."""
    lo, hi = shard_bounds(sky.Ncomponents, world_size, rank)
    inds = torch.arange(lo, hi, dtype=torch.int64, device=sky.device)
    return sky.select(component_inds=inds, inplace=False)


def exchange_counts(n_local, device=None, group=None):
    """All-gather the per-rank above-horizon counts.

    Returns ``(counts, col0, total)``: counts per rank (list of int), this rank's column offset in the
    global compacted array and the global number of above-horizon components.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return [int(n_local)], 0, int(n_local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    parts = [torch.empty(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    counts = [int(c.item()) for c in parts]
    return counts, int(sum(counts[:rank])), int(sum(counts))


def gather_columns(local, counts, group=None):
    """Assemble the global compacted array from per-rank column blocks.

    ``local``: (..., n_r) tensor of this rank; ``counts``: list from :func:`exchange_counts`.  Returns the
    (..., sum counts) tensor on every rank, ranks' blocks in rank order (= catalogue order).
    """
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    lead = tuple(local.shape[:-1])
    nmax = max(counts) if counts else 0
    total = int(sum(counts))
    if nmax == 0:
        return local.new_empty(lead + (0,))
    # real view so that gloo (no complex support) and NCCL share one code path
    is_complex = local.is_complex()
    loc = torch.view_as_real(local.contiguous()) if is_complex else local.contiguous()
    tail = (2,) if is_complex else ()
    ax = len(lead)  # the column axis
    padded = loc.new_zeros(lead + (nmax,) + tail)
    padded.narrow(ax, 0, loc.shape[ax]).copy_(loc)
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    full = loc.new_empty(lead + (total,) + tail)
    off = 0
    for r, c in enumerate(counts):
        if c:
            full.narrow(ax, off, c).copy_(gathered[r].narrow(ax, 0, c))
        off += c
    return torch.view_as_complex(full) if is_complex else full


class PeerGather:
    """Gathered coherency array in symmetric memory, filled by the kernels themselves over NVLink.

    ``PeerGather(nfreqs, max_cols)`` allocates, once, a ``(2, 2, nfreqs, ld)`` complex128 buffer on every rank
    (``ld`` = ``max_cols`` rounded up to 16 columns: 256-byte rows; ``max_cols`` must cover the gathered columns plus
    the largest rank's component count, because every launch is sized for its rank's whole block) with
    ``torch.distributed._symmetric_memory`` and maps every peer's copy.  :meth:`coherency_calc` then runs, per
    time step: the 8-byte count exchange (column offsets), one ``rsk_coherency`` launch per destination whose
    ``out`` pointer is the *destination's* buffer and whose ``col0`` is this rank's offset, and a device-side
    barrier.  ``dst=None`` leaves the full array on every rank (all-gather semantics: each rank stores its block
    world_size times, once locally and world_size - 1 times through NVLink); ``dst=r`` assembles it on rank r
    only.  No NCCL data collective is involved and nothing is staged: the compute kernel's 128-bit streaming
    stores are the transfer.
    """

    def __init__(self, nfreqs, max_cols, device=None, group=None, streams=2):
        import torch.distributed._symmetric_memory as symm_mem

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGather needs an initialised torch.distributed process group (nccl)")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.nfreqs = int(nfreqs)
        self.ld = max(16, -(-int(max_cols) // 16) * 16)
        shape = (2, 2, self.nfreqs, self.ld)
        self.buf = symm_mem.empty(*shape, dtype=torch.complex128, device=self.device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.peers = [self.buf if r == self.rank else self.handle.get_buffer(r, shape, torch.complex128)
                      for r in range(self.world)]
        # launches for different destinations go round-robin over a few streams so that several NVLink
        # destinations are in flight at once (one remote-store kernel alone reaches ~520 of the 900 GB/s)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(max(1, int(streams)))]

    def coherency_calc(self, sky, freqs=None, dst=None, route="copy", slabs=4, **kwargs):
        """Sharded ``sky.coherency_calc`` assembled in the gathered array.

        ``route="copy"`` (default): this rank's columns are computed into its own copy of the array in ``slabs``
        column slabs; as soon as a slab is complete, the copy engines push it into the same place of every
        destination's array (``rsk_copy_2d``: one pitched peer-to-peer copy per slab and destination, on side
        streams), so the NVLink transfer of slab k overlaps the kernel of slab k + 1.  Measured on B200 (bench.py
        ``gather`` record): ~0.9 of the neighbour-to-neighbour copy rate.
        ``route="store"``: one ``rsk_coherency`` launch per destination whose ``out`` is the destination's buffer -- the
        kernel's own 128-bit stores travel over NVLink (0.64-0.72 of the copy rate: a remote-store kernel keeps fewer
        bytes in flight than the copy engines).

        Returns ``(view, counts)``: the (2, 2, nfreqs, sum N_above) view of this rank's buffer (valid on the
        destination rank(s) once the call returns) and the per-rank above-horizon counts."""
        if route == "copy":
            return self._coherency_calc_copy(sky, freqs, dst, slabs, **kwargs)
        if route != "store":
            raise ValueError("route must be 'copy' or 'store'")
        _, n_above = sky.compact_above_horizon()
        counts, col0, total = exchange_counts(n_above, self.device, self.group)
        if total > self.ld:
            raise ValueError(f"gathered array needs {total} columns, the buffer has {self.ld}")
        if col0 + sky.Ncomponents > self.ld:
            # the launches below leave the count on the device, so each is sized for the rank's whole block
            raise ValueError(f"PeerGather needs max_cols >= column offset + Ncomponents of every rank "
                             f"({col0} + {sky.Ncomponents} > {self.ld}); allocate sum(N_above) + max(Ncomponents)")
        # all-gather: at step k rank i writes to rank (i + k) mod world, so every step is a permutation -- each GPU
        # receives from exactly one sender at a time instead of all ranks converging on the same destination
        targets = [(self.rank + k) % self.world for k in range(self.world)] if dst is None else [int(dst)]
        self.handle.barrier()  # nobody is still reading the previous step's array
        main = torch.cuda.current_stream(self.device)
        if len(targets) == 1 or len(self.streams) == 1:
            for r in targets:
                sky.coherency_calc(freqs=freqs, out=self.peers[r], ld_out=self.ld, col0=col0, sync=False, **kwargs)
        else:
            for st in self.streams:
                st.wait_stream(main)
            for k, r in enumerate(targets):
                with torch.cuda.stream(self.streams[k % len(self.streams)]):
                    sky.coherency_calc(freqs=freqs, out=self.peers[r], ld_out=self.ld, col0=col0, sync=False, **kwargs)
            for st in self.streams:
                main.wait_stream(st)
        self.handle.barrier()  # every rank's stores have landed
        return self.buf[..., :total], counts

    def _coherency_calc_copy(self, sky, freqs, dst, slabs, freq_interp_kind="cubic", nan_handling="clip"):
        import ctypes

        from . import _lib

        plan = sky._plan_coherency(freqs, freq_interp_kind, need_count=True, nan_handling=nan_handling)
        n_above, nf = plan["n_above"], plan["nf"]
        if nf != self.nfreqs:
            raise ValueError(f"the buffer holds {self.nfreqs} channels, the request has {nf}")
        counts, col0, total = exchange_counts(n_above, self.device, self.group)
        if total > self.ld:
            raise ValueError(f"gathered array needs {total} columns, the buffer has {self.ld}")
        # every rank computes its own columns into its own copy; the destinations other than itself get them by copy
        targets = [(self.rank + k) % self.world for k in range(1, self.world)] if dst is None else (
            [] if int(dst) == self.rank else [int(dst)])
        self.handle.barrier()  # nobody is still reading the previous step's array
        main = torch.cuda.current_stream(self.device)
        lib = _lib.load()
        # slabs of at least 32k columns (0.5 MB rows): measured on 8 B200 (profiles/r02_gather_sweep_8gpu.txt), shorter
        # rows cost more in copy-engine efficiency than the overlap with the next slab's kernel gains
        nslab = max(1, min(int(slabs), n_above // 32768)) if n_above > 0 else 0
        per = -(-(-(-n_above // max(nslab, 1))) // 16) * 16 if nslab else 0  # whole 256-byte groups per slab
        esize = 16
        base = self.buf.data_ptr()
        k = 0
        for s in range(nslab):
            c0, c1 = s * per, min(n_above, (s + 1) * per)
            if c0 >= c1:
                break
            sub = dict(plan)
            sub["idx"] = None if plan["idx"] is None else plan["idx"][c0:c1]
            sub["n_above"], sub["count_dev"] = c1 - c0, None
            if plan["idx"] is None:  # no horizon mask: the slab is the component range itself
                sub["idx"] = torch.arange(c0, c1, dtype=torch.int64, device=self.device)
            sky._launch_coherency(sub, 0, nf, self.buf, self.ld, col0 + c0)
            done = torch.cuda.Event()
            done.record(main)
            off = (col0 + c0) * esize
            for r in targets:
                st = self.streams[k % len(self.streams)]
                k += 1
                st.wait_event(done)
                with torch.cuda.device(self.device):
                    code = lib.rsk_copy_2d(ctypes.c_void_p(self.peers[r].data_ptr() + off), self.ld * esize,
                                           ctypes.c_void_p(base + off), self.ld * esize, (c1 - c0) * esize, 4 * nf,
                                           ctypes.c_void_p(st.cuda_stream))
                _lib.check("rsk_copy_2d", code)
        for st in self.streams:
            main.wait_stream(st)
        self.handle.barrier()  # every rank's copies have landed
        return self.buf[..., :total], counts
