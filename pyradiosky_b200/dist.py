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
