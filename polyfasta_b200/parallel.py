"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink), work split by columns or by loci.

Two ways the path shards:
  * within one alignment: contiguous, codon-aligned column ranges per rank; every rank scans its range and the int64
    vectors ([S, H, SFS...] per population, or the PFA_CDS_LEN codon accumulators) are summed with ONE all-reduce --
    integer sums, so the result is bit-identical for any number of ranks;
  * across loci (--dir): locus i (in the reference's sorted() order, PolyFastA.py:104) goes to rank i mod world, no
    collective; rank 0 gathers the finished rows and prints them in sorted order.
"""


def shard_columns(total, world, rank, multiple=3):
    """[begin, end) of rank's column range: boundaries are multiples of `multiple` (codon-aligned so that --cds shards
    never split a codon); the last rank takes the remainder, including a trailing partial codon."""
    per = (total // world) // multiple * multiple
    begin = rank * per
    end = total if rank == world - 1 else (rank + 1) * per
    return begin, end


def plan_units(paths, sizes, shard_min_bytes, batch_files, batch_bytes):
    """the work units of a multi-rank run over `paths` (already in the reference's sorted() order, PolyFastA.py:104):
    ("all", [path]) for a file of at least shard_min_bytes -- ONE alignment whose columns are split over all ranks -- and
    ("chunk", [paths...]) for runs of smaller files (at most batch_files files / batch_bytes bytes each), which go to
    single ranks round-robin (chunk_owner).  Units keep the order of `paths`."""
    units, cur, cur_bytes = [], [], 0
    for p, sz in zip(paths, sizes):
        if shard_min_bytes is not None and sz >= shard_min_bytes:
            if cur:
                units.append(("chunk", cur))
                cur, cur_bytes = [], 0
            units.append(("all", [p]))
            continue
        if cur and (len(cur) >= batch_files or cur_bytes + sz > batch_bytes):
            units.append(("chunk", cur))
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += sz
    if cur:
        units.append(("chunk", cur))
    return units


def chunk_owner(j, world):
    """the rank (or device slot) that processes the j-th chunk: round-robin, no collective"""
    return j % world


def bind_near_gpu(device):
    """best effort: keep this process (and the ingest threads it starts) on the CPUs of the GPU's NUMA node, so that the pinned
    staging buffers it allocates and packs are local to the PCIe root the uploads leave through.  One process per GPU on a
    two-socket host otherwise reads half of its text across the socket link.  POLYFASTA_NUMA_BIND=0 turns it off.
    Returns the CPU set now in force (or None when nothing was changed)."""
    import os
    if os.environ.get("POLYFASTA_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cur = os.sched_getaffinity(0)
        new = cur & cpus
        if not new or new == cur:
            return None
        os.sched_setaffinity(0, new)
        return new
    except Exception:
        return None


def connect_exchange(ctx, cap_words, group=None):
    """one api.Exchange per rank, connected over the ranks of `group`: the CUDA IPC handles of the symmetric buffers travel
    through all_gather_object (host plumbing); afterwards Alignment.site_stats_xchg / cds_stats_xchg sum the shard vectors
    inside the scan kernel over NVLink, without a collective call.  Collective: every rank of the group must call it."""
    import torch.distributed as dist
    from . import api
    x = api.Exchange(ctx, cap_words)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        x.connect(0, 1, [bytes(api.Exchange.HANDLE_BYTES)])   # one rank: nothing to map
        return x
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handles = [None] * world
    dist.all_gather_object(handles, x.export(), group=group)
    x.connect(rank, world, handles)
    dist.barrier(group=group)  # every rank has mapped every buffer before the first kernel touches one
    return x


def gather_rows(rows):
    """--dir mode: every rank contributes [(locus_index, text), ...]; rank 0 gets them back merged in index order"""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sorted(rows)
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(rows, out, dst=0)
    if dist.get_rank() != 0:
        return None
    return sorted(r for part in out for r in part)
