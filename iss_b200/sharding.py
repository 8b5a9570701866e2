"""Event sharding across the GPUs of one node.

Events are independent given the yields and every random stream is keyed by the global event index,
so rank r of N simply samples its own event range with the same seed: no data-path collective, and
the union of the ranks' outputs is bit-identical to a single-GPU run.  Only the QA block (plain
sums, include/iss_cuda.h) is reduced, with one all-reduce (NCCL on GPUs; gloo in the CPU tests).

The smooth-spectra integrator (iss_cuda_spectra) shards the other way: every rank holds the whole
surface and integrates its own species; the [npT][nphi] tables are gathered, no reduction at all
(the sum over cells stays on one GPU, in a chunk order that does not depend on the rank count).

Surface-chunk sharding (the alternative for surfaces like C5): every rank
holds a contiguous cell range and samples, for ALL events, the hadrons whose cell lies in it.  The
one collective is an all-gather of the per-tile (1024 cells) yield sums, [nspecies][ntile] doubles;
every rank then combines them in the same fixed order (include/iss_cuda.h,
iss_cuda_chunk_yields_finish), so totals, multiplicities and chosen cells are bit-identical to a
single-GPU run and the union of the ranks' lists is that run's list.
"""
import torch
import torch.distributed as dist

CHUNK_ALIGN = 4096      # ISS_CHUNK_ALIGN: one node of level 2 of the cell-search tree
TILE = 1024             # cells per tile of the fixed-order yield sums


def split_cells(ncell, world):
    """Contiguous cell ranges [b, e) per rank, boundaries on multiples of CHUNK_ALIGN, sizes as
    equal as the alignment allows; trailing ranks may be empty for small surfaces."""
    ncell, world = int(ncell), int(world)
    nblk = (ncell + CHUNK_ALIGN - 1)//CHUNK_ALIGN
    base, rem = divmod(nblk, world)
    out, b = [], 0
    for r in range(world):
        e = b + (base + (1 if r < rem else 0))*CHUNK_ALIGN
        out.append((min(b, ncell), min(e, ncell)))
        b = e
    return out


def split_cells_weighted(block_cost, ncell, world):
    """Contiguous cell ranges [b, e) per rank on multiples of CHUNK_ALIGN whose summed cost is as even
    as a cut at block boundaries allows.  block_cost[j]: cost of cells [4096 j, 4096 (j+1)), e.g.
    a*cells + b*events*yield from iss_cuda_chunk_block_yields; every rank gets at least one block."""
    import numpy as np
    cost = np.asarray(block_cost, dtype=np.float64)
    nblk = (int(ncell) + CHUNK_ALIGN - 1)//CHUNK_ALIGN
    if len(cost) != nblk:
        raise ValueError("block_cost must have ceil(ncell/4096) = %d entries" % nblk)
    if nblk < world:
        return split_cells(ncell, world)
    cum = np.concatenate([[0.], np.cumsum(cost)])
    cuts = [0]
    for r in range(1, world):
        # first boundary whose cumulative cost reaches r/world of the total, leaving a block for
        # every rank before and after
        j = int(np.searchsorted(cum, cum[-1]*r/world, side="left"))
        if j > 0 and abs(cum[j - 1] - cum[-1]*r/world) < abs(cum[j] - cum[-1]*r/world):
            j -= 1
        j = max(j, cuts[-1] + 1)
        j = min(j, nblk - (world - r))
        cuts.append(j)
    cuts.append(nblk)
    return [(min(cuts[r]*CHUNK_ALIGN, int(ncell)), min(cuts[r + 1]*CHUNK_ALIGN, int(ncell)))
            for r in range(world)]


def ntiles_of(cell_range):
    return (cell_range[1] - cell_range[0] + TILE - 1)//TILE


def gather_tile_sums(local, ranges, nspecies):
    """local: float64 tensor [nspecies, ntiles_of(ranges[rank])] (device of the backend: CUDA for
    NCCL, CPU for gloo) -> list over ranks of contiguous [nspecies, ntiles_of(ranges[r])] tensors.
    One all_gather of equally padded blocks."""
    world = len(ranges)
    nt = [ntiles_of(r) for r in ranges]
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return [local.contiguous()]
    width = max(nt)
    pad = torch.zeros((int(nspecies), width), dtype=torch.float64, device=local.device)
    pad[:, :local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return [parts[r][:, :nt[r]].contiguous() for r in range(world)]


def chunk_yields(engine, cells, rank, world, device=None):
    """Surface-chunk yields of one rank: uploads cells[b:e] of the whole surface `cells`
    (float32 [ncell, 28]), runs the local part, all-gathers the tile sums over the process group
    (NCCL) and finishes.  Returns (dN per species over the whole surface, (b, e))."""
    ranges = split_cells(len(cells), world)
    b, e = ranges[rank]
    if e <= b:
        raise ValueError("rank %d has no cells: use fewer ranks for a surface of %d cells" %
                         (rank, len(cells)))
    engine.upload_surface(cells[b:e])
    engine.set_surface_chunk(b, len(cells))
    ptr, nt = engine.chunk_yields_local()
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    local = device_block_as_tensor(ptr, engine.nspecies*nt, dev).view(engine.nspecies, nt)
    blocks = gather_tile_sums(local, ranges, engine.nspecies)
    torch.cuda.synchronize()
    dN = engine.chunk_yields_finish([t.data_ptr() for t in blocks], [t.shape[1] for t in blocks],
                                    on_device=True)
    return dN, (b, e)


def split_events(nev_total, world):
    """Contiguous, disjoint ranges covering [0, nev_total): rank r gets [b[r], b[r+1])."""
    base, rem = divmod(int(nev_total), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def weak_event_range(step, rank, world, events_per_rank):
    """Weak-scaling benchmark layout: every rank samples `events_per_rank` new events per step."""
    begin = (int(step)*int(world) + int(rank))*int(events_per_rank)
    return begin, begin + int(events_per_rank)


def split_species(nspecies, world):
    """Round-robin species lists per rank (species are mass-sorted and the cost per species is
    uniform, so round-robin balances any list length): rank r gets indices r, r + N, ..."""
    return [list(range(r, int(nspecies), int(world))) for r in range(int(world))]


def gather_species_tables(local, nspecies, world, rank):
    """local: float64 tensor [len(split_species(...)[rank]), npT, nphi] -> [nspecies, npT, nphi] on
    every rank (all_gather of equally padded blocks, then un-interleaved)."""
    per = (int(nspecies) + world - 1)//world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    if world > 1:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
    else:
        parts = [pad]
    out = torch.zeros((int(nspecies),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, idx in enumerate(split_species(nspecies, world)):
        out[idx] = parts[r][:len(idx)]
    return out


def allreduce_sum_(t):
    """In-place sum over ranks of a tensor (QA block, counters); no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def join_engine_communicator(engine):
    """Gives the engine's handle its own NCCL communicator over the ranks of the torch process group
    (iss_cuda_nccl_unique_id on rank 0, the 128-byte id broadcast with torch.distributed,
    iss_cuda_nccl_init on every rank): afterwards iss_cuda_histograms_allreduce(h, NULL) reduces the
    QA block through the C ABI, the way a C++ host does it.  Returns True on success."""
    import ctypes as C
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return False
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = torch.zeros(128, dtype=torch.uint8)
    ok = torch.ones(1, dtype=torch.int32)
    if rank == 0:
        buf = (C.c_ubyte*128)()
        if engine.L.iss_cuda_nccl_unique_id(buf) == 0:
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        else:
            ok[0] = 0
    dev = torch.device("cuda", torch.cuda.current_device())
    ident, ok = ident.to(dev), ok.to(dev)
    dist.broadcast(ident, 0)
    dist.broadcast(ok, 0)
    if int(ok.item()) == 0:
        return False
    raw = (C.c_ubyte*128)(*[int(x) for x in ident.cpu().tolist()])
    rc = torch.tensor([engine.L.iss_cuda_nccl_init(engine.h, raw, rank, world)], dtype=torch.int32, device=dev)
    dist.all_reduce(rc, op=dist.ReduceOp.MAX)
    return int(rc.item()) == 0


def device_block_as_tensor(device_ptr, n_doubles, device):
    """Wraps a device pointer of the engine (e.g. iss_cuda_qa_device_ptr) as a torch tensor so that
    NCCL can reduce it in place."""
    class _Ext:
        pass
    o = _Ext()
    o.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8",
                                  "data": (int(device_ptr), False), "version": 2}
    return torch.as_tensor(o, device=device)
