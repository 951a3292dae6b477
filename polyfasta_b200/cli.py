"""Drop-in command line of PolyFastA.py on top of the B200 library.

Flags, defaults, CSV columns, screen text, `#` data-error rows and exit codes follow the reference's main() and
print_result() (PolyFastA.py:13-225); all counting and the fp64 statistics run on the GPU.  `-p/--pops` implements
the semantics the reference intends (PolyFastA.py:123-134: header-substring match, one row per key in the given
order); the reference itself raises TypeError there under Python 3 (:126)."""
import argparse
import os
import sys

from . import api

DESCRIPTION = """
    Fast estimator of nucleotide diversity (theta_pi),
    Watterson's theta (theta_w), and Tajimas\'s D for coding and
    non-coding sequences\n"""

EPILOG = """

    It can also print a vector of the folded site frequency spectrum and
    correct theta_pi for multiple hits (jukes-cantor).

    Examples:
    python PolyFastA.py -f myAlignment.fas -p pop1,pop2
    \"-p pop1,pop2\" assumes that the alignment has sequences that are labeled:

    >pop1_ind1_XXX
    ATGC...
    >pop1_ind2_XXX
    ATGC...
    >pop2_ind1_XXX
    ATGC...
    >pop2_ind2_XXX

    Alignment is inframe cooding sequence:
    python PolyFastA.py -f myAlignment.fas -p pop1,pop2 --cds

    Want Jukes-Cantor corrected estimates for CDS:
    python PolyFastA.py -f myAlignment.fas -p pop1,pop2 --cds --jc

    FASTA files in directory:
    python PolyFastA.py -d myFastaDir/ --out allpoly.csv

    Read from pipe:
    python PolyFastA.py --pipe -p pop1,pop2

    The format is not strict but the identifier (e.g. pop1) needs to be somewhere in the header.\n"""

# (long, short, kwargs): the flag table of PolyFastA.py:51-77
FLAGS = [
    ("--file", "-f", dict(nargs="*", type=str, default="", help="one or several alignment files in FASTA format.")),
    ("--dir", "-d", dict(type=str, default=".", help="directory containing FASTA files only.")),
    ("--pops", "-p", dict(nargs="?", default=False, metavar="pop1,pop2,pop3", type=str,
                          help="split alignment by populations. A comma-separated list of strings that are found in the sequence headers.")),
    ("--out", "-o", dict(type=str, default="", help="name of output file. (default/empty will print to screen")),
    ("--pipe", "-i", dict(action="store_true", default=False, help="if FASTA file is being piped in from STDIN.")),
    ("--cds", "-c", dict(action="store_true", default=False,
                         help="the alignment is protein coding. Will split into synonymous and nonsynonymous sites.")),
    ("--silent", "-s", dict(action="store_true", default=False, help="suppress header and verbose output.")),
    ("--name", "-n", dict(type=str, default=False, help="name of the DNA region to show in output.")),
    ("--jc", None, dict(action="store_true", default=False, help="Jukes-Cantor correction for Pi.")),
]

HEADER_NONCDS = "file,seqlen,pop,N,seg_sites,pi,theta,tajimasD"
HEADER_CDS = ("file,sites_S,sites_N,pop,N,seg_sites_S,seg_sites_N,pi_S,pi_N,theta_S,theta_N,"
              "tajimasD_S,tajimasD_N,nstops")


def build_parser():
    parser = argparse.ArgumentParser(prog="PolyFastA.py", formatter_class=argparse.RawTextHelpFormatter,
                                     description=DESCRIPTION, epilog=EPILOG)
    for long, short, kw in FLAGS:
        names = [long] if short is None else [long, short]
        parser.add_argument(*names, **kw)
    return parser


class Sink:
    """where rows go: the screen, or --out (header truncates, everything else appends, plus the \\r progress text)"""

    def __init__(self, out, silent):
        self.out = out
        self.silent = silent

    def header(self, cds):
        line = HEADER_CDS if cds else HEADER_NONCDS
        if self.out:
            with open(self.out, "w") as o:
                o.write(line + "\n")
        else:
            print(line)

    def note(self, line):
        """`#` data-problem rows (PolyFastA.py:115-119,127-131,136-140)"""
        if self.out:
            with open(self.out, "a") as o:
                o.write(line + "\n")
        else:
            print(line)

    def row(self, line, file, pop):
        if self.out:
            with open(self.out, "a") as o:
                if not self.silent:
                    print("\r", f"Writting to {self.out}, pop: {pop:<10s}, parsing: {file:<15s}", end="", flush=True)
                o.write(line + "\n")
        else:
            print(line)


def format_rows(ctx, file, seqlen, cds, jc, pops, site, cdsst):
    """rows of print_result.no_header (PolyFastA.py:148-198) for the populations of one alignment.
    pops: list of (label, n); site: K2 results per population; cdsst: K4 results per population or None."""
    todo = []
    for q, (label, n) in enumerate(pops):
        if cds:
            c = cdsst[q]
            ssites = c["ssites"]
            nsites = (seqlen - c["missing"]) - ssites
            if site[q]["S"]:
                todo.append((n, c["S_s"], c["H_s"], ssites, jc))
                todo.append((n, c["S_n"], c["H_n"], nsites, jc))
        elif site[q]["S"]:
            todo.append((n, site[q]["S"], site[q]["H"], seqlen, jc))
    fin = iter(ctx.finalize(todo))
    lines = []
    for q, (label, n) in enumerate(pops):
        if cds:
            c = cdsst[q]
            ssites = c["ssites"]
            nsites = (seqlen - c["missing"]) - ssites
            head = f"{file},{round(ssites, 2)},{round(nsites, 2)},{label},{n}"
            if site[q]["S"] == 0:
                lines.append(head + ",0,0,0,NA,0,0,0,NA,0")                                   # :160/:162
            else:
                s, m = next(fin), next(fin)
                lines.append(head + f",{s[0]},{m[0]},{s[1]},{m[1]},{s[2]},{m[2]},{s[3]},{m[3]},{c['nstops']}")  # :178/:180
        elif site[q]["S"] == 0:
            lines.append(f"{file},{seqlen},{label},{n},0,0,0,NA")                              # :187/:189
        else:
            r = next(fin)
            lines.append(f"{file},{seqlen},{label},{n},{r[0]},{r[1]},{r[2]},{r[3]}")           # :196/:198
    return lines


class RankGroup:
    """the ranks of a torchrun launch (one process per GPU): rendezvous and row gathering go through torch.distributed
    (gloo, host plumbing only); the per-shard count vectors are summed inside the scan kernels over NVLink (api.Exchange)"""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self.host_cpus = os.cpu_count() or 1
        from . import parallel
        parallel.bind_near_gpu(self.device)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
        self.xchg = None
        # the in-kernel exchange maps every rank's buffer with CUDA IPC: one host, at most 16 ranks.  Anything else (several
        # nodes, more ranks) sums the small count vectors on the host through the rendezvous backend instead.
        import socket
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname())
        self.fused = self.world <= 16 and len(set(hosts)) == 1 and os.environ.get("POLYFASTA_HOST_ALLREDUCE", "0") != "1"

    def parse_once(self, path, threads):
        """a large file every rank needs: rank 0 parses it with all host threads and hands the others the row layout (a few bytes
        per row); they map the file and adopt it.  Files that were not mapped in place are parsed by every rank."""
        blob = [None]
        fasta = None
        if self.rank == 0:
            fasta = api.parse_files([path], threads=max(threads, (os.cpu_count() or 1)))[0]
            blob = [fasta.export_layout() if isinstance(fasta, api.Fasta) else None]
        self.dist.broadcast_object_list(blob, src=0)
        if self.rank == 0:
            return fasta
        if blob[0] is None:
            return api.parse_files([path], threads=threads)[0]
        return api.Fasta.from_layout(path, blob[0])

    def exchange(self, ctx, words):
        """an exchange of at least `words` int64 (collective: every rank asks for the same size at the same point)"""
        from . import parallel
        if self.xchg is None or self.xchg.capacity < words:
            if self.xchg is not None:
                ctx.sync()
                self.dist.barrier()
                self.xchg.close()
            self.xchg = parallel.connect_exchange(ctx, max(words, 1 << 14))
        return self.xchg

    def close(self, ctx):
        if self.xchg is not None:
            ctx.sync()
            self.dist.barrier()
            self.xchg.close()
            self.xchg = None


def sharded_alignment_rows(ctx, group, fasta, file, cds, jc, found):
    """rows of one large alignment whose COLUMNS are split over the ranks: every rank uploads and scans its
    codon-aligned range, the scan kernels sum the integer vectors over all ranks, rank 0 formats the rows"""
    import numpy as np
    import torch
    from . import parallel
    c0, c1 = parallel.shard_columns(fasta.seqlen, group.world, group.rank)
    aln = api.Alignment.from_fasta(ctx, fasta, c0, c1)
    try:
        if found[0][1] is not None:
            m = np.stack([mask for _, mask, _ in found])
            api.check(api.lib().pfa_aln_set_pops(aln.handle, m.ctypes.data, len(found)), ctx.handle)
        k = len(found)
        ns, nc = aln.site_len(), api.PFA_CDS_LEN * k
        d_site = torch.zeros(ns, dtype=torch.int64, device="cuda:%d" % ctx.device)
        d_cds = torch.zeros(nc, dtype=torch.int64, device="cuda:%d" % ctx.device)
        torch.cuda.synchronize(ctx.device)
        if group.fused:
            x = group.exchange(ctx, ns + nc)
            # the ranks meet on the host AFTER their uploads (parse + ingest times differ by seconds on cold files), so the
            # wait inside the kernels only has to cover launch skew
            ctx.sync()
            group.dist.barrier()
            if cds:   # K2 leaves its vector in the exchange's buffer, K4's epilogue pushes both: ONE exchange per alignment
                d_both = torch.zeros(ns + nc, dtype=torch.int64, device="cuda:%d" % ctx.device)
                torch.cuda.synchronize(ctx.device)
                aln.site_cds_stats_xchg(x, d_both.data_ptr())
            else:
                aln.site_stats_xchg(x, d_site.data_ptr())
            ctx.sync()
            if x.timed_out():
                raise api.PolyFastaError(1, "a rank did not arrive at the exchange")
            if cds:
                h_both = d_both.cpu()
                h_site, h_cds = h_both[:ns], h_both[ns:]
            else:
                h_site, h_cds = d_site.cpu(), d_cds.cpu()
        else:
            aln.site_stats_device(d_site.data_ptr())
            if cds:
                aln.cds_stats_device(d_cds.data_ptr())
            ctx.sync()
            h_site, h_cds = d_site.cpu(), d_cds.cpu()
            group.dist.all_reduce(h_site)
            if cds:
                group.dist.all_reduce(h_cds)
        site = aln.unpack_site(h_site.numpy())
        cdsst = None
        if cds:
            raw = h_cds.numpy().reshape(k, api.PFA_CDS_LEN)
            ss = ctx.cds_ssites(raw)
            cdsst = []
            for q in range(k):
                d = aln.unpack_cds(raw[q])
                d["ssites"] = float(ss[q])
                cdsst.append(d)
    finally:
        aln.free()
    if group.rank != 0:
        return [None] * len(found)
    return format_rows(ctx, file, fasta.seqlen, cds, jc, [(label, rows) for label, _, rows in found], site, cdsst)


BATCH_FILES = int(os.environ.get("POLYFASTA_BATCH_FILES", "512"))   # loci per batched GPU pass
SHARD_MIN_BYTES = int(os.environ.get("POLYFASTA_SHARD_MIN_BYTES", str(64 << 20)))   # files this large are column-sharded over the ranks
BATCH_BYTES = 1 << 30          # ... or this much text


def plan_populations(fasta, popkeys):
    """[(label, mask | None, rows)]: header-substring split (PolyFastA.py:123-125); no keys = one population 'NA'"""
    if popkeys is None:
        return [("NA", None, fasta.nseq)]
    plan = []
    for key in popkeys:
        mask, hits = api.match_mask(fasta, key)
        plan.append((key, mask, hits))
    return plan


def single_alignment_rows(ctx, fasta, file, cds, jc, found):
    """rows of one alignment through the single-alignment kernels (large alignments, and --cds)"""
    import numpy as np
    aln = api.Alignment.from_fasta(ctx, fasta)
    try:
        if found[0][1] is not None:
            m = np.stack([mask for _, mask, _ in found])
            api.check(api.lib().pfa_aln_set_pops(aln.handle, m.ctypes.data, len(found)), ctx.handle)
        site = aln.site_stats()
        cdsst = aln.cds_stats() if cds else None
    finally:
        aln.free()
    return format_rows(ctx, file, fasta.seqlen, cds, jc, [(label, rows) for label, _, rows in found], site, cdsst)


def process_chunk(ctx, batch, items, cds, jc, popkeys, group=None):
    """items: [(path, Fasta | exception)] in output order -> ordered actions ('stdout' | 'note' | 'row', ...).
    Small non-CDS loci of the chunk share ONE batched GPU pass; everything else goes through the single-alignment path."""
    actions = []
    pending = []   # (position in actions, locus, pop, file, seqlen, label, n)
    own_batch = None
    if batch is not None:
        batch.clear()
    for path, fasta in items:
        file = path.split("/")[-1]
        if isinstance(fasta, api.NotFasta):
            actions.append(("stdout", f"# file {path} is not FASTA!"))                         # PolyFastA.py:247
            continue
        if isinstance(fasta, Exception):
            actions.append(("raise", fasta))
            break
        if fasta.seqlen < 0:
            actions.append(("note", f"# Sequences do not have the same length: {file}"))        # :135-140
            continue
        if cds and fasta.seqlen % 3 != 0:
            actions.append(("note", f"# CDS sequence length is not a multiple of 3: {file}"))   # :114-119
        plan = plan_populations(fasta, popkeys)
        found = [p for p in plan if p[2] > 0]
        rows = iter(())
        if found and group is not None:
            rows = iter(sharded_alignment_rows(ctx, group, fasta, file, cds, jc, found))
        elif found and (cds or not api.Batch.fits(fasta)):
            rows = iter(single_alignment_rows(ctx, fasta, file, cds, jc, found))
        elif found:
            import numpy as np
            masks = None if found[0][1] is None else np.stack([m for _, m, _ in found])
            if batch is None:   # a caller that expected only large alignments: stage the small one in a batch of our own
                batch = own_batch = api.Batch(ctx)
            locus = batch.add(fasta, masks)
        q = 0
        for label, mask, hits in plan:
            if hits == 0:
                actions.append(("note", f"# Pop {label} string was not found in fasta headers."))   # :126-132
            elif group is not None or cds or not api.Batch.fits(fasta):
                actions.append(("row", next(rows), file, label))
            else:
                pending.append((len(actions), locus, q, file, fasta.seqlen, label, hits))
                actions.append(None)
                q += 1
    if pending:
        batch.run(jc)
        for pos, locus, q, file, seqlen, label, n in pending:
            r = batch.result(locus, q)
            if r["S"] == 0:
                line = f"{file},{seqlen},{label},{n},0,0,0,NA"                                     # :187/:189
            else:
                p = r["poly"]
                line = f"{file},{seqlen},{label},{n},{p[0]},{p[1]},{p[2]},{p[3]}"                  # :196/:198
            actions[pos] = ("row", line, file, label)
    for _, fasta in items:
        if isinstance(fasta, api.Fasta):
            fasta.close()
    if own_batch is not None:
        own_batch.close()
    return actions


def noncds_line(file, seqlen, label, n, r):
    if r["S"] == 0:
        return f"{file},{seqlen},{label},{n},0,0,0,NA"                                         # PolyFastA.py:187/:189
    p = r["poly"]
    return f"{file},{seqlen},{label},{n},{p[0]},{p[1]},{p[2]},{p[3]}"                          # :196/:198


def cds_line(file, seqlen, label, n, site, c):
    """the --cds row of print_result.no_header (PolyFastA.py:160-180) from the batched results of one (locus, population)"""
    ssites = c["ssites"]
    nsites = (seqlen - c["missing"]) - ssites
    head = f"{file},{round(ssites, 2)},{round(nsites, 2)},{label},{n}"
    if site["S"] == 0:
        return head + ",0,0,0,NA,0,0,0,NA,0"                                                   # :160/:162
    s, m = c["poly_s"], c["poly_n"]
    return head + f",{s[0]},{m[0]},{s[1]},{m[1]},{s[2]},{m[2]},{s[3]},{m[3]},{c['nstops']}"     # :178/:180


def process_chunk_native(ctx, batch, paths, jc, popkeys, threads, cds=False):
    """--dir chunk: ONE native call reads, parses, splits and stages every file (host threads), ONE batched GPU pass computes
    every (locus, population) -- the site scan and, with --cds, the codon scan -- then the rows are formatted in order"""
    from ._lib import PFA_BATCH_TOO_BIG, PFA_ERR_IO, PFA_ERR_NON_ASCII, PFA_ERR_NOT_FASTA, PFA_ERR_RAGGED, PFA_OK
    import time
    t0 = time.perf_counter()
    batch.clear()
    # files too large for the batched kernels skip the native staging (it would read and parse them only to refuse them)
    def _size(p):
        try:
            return os.path.getsize(p)
        except OSError:
            return 0
    big = [_size(p) > api.Batch.MAX_LOCUS_BYTES for p in paths]
    small_info = iter(batch.add_files([p for p, b in zip(paths, big) if not b], popkeys or (), threads))
    info = [{"status": PFA_BATCH_TOO_BIG} if b else next(small_info) for b in big]
    t1 = time.perf_counter()
    labels = popkeys if popkeys is not None else ["NA"]
    actions, pending = [], []
    for path, fi in zip(paths, info):
        file = path.split("/")[-1]
        st = fi["status"]
        if st == PFA_ERR_NOT_FASTA:
            actions.append(("stdout", f"# file {path} is not FASTA!"))
        elif st == PFA_ERR_IO:
            actions.append(("raise", OSError("cannot read %s" % path)))
            break
        elif st == PFA_ERR_NON_ASCII:
            actions.append(("raise", ValueError("%s: non-ASCII bytes in sequence lines are not supported" % path)))
            break
        elif st == PFA_ERR_RAGGED:
            actions.append(("note", f"# Sequences do not have the same length: {file}"))
        elif st == PFA_BATCH_TOO_BIG:
            actions.extend(process_chunk(ctx, None, [(path, api.parse_files([path], threads=threads)[0])], cds, jc, popkeys))
        elif st == PFA_OK:
            if cds and fi["L"] % 3 != 0:
                actions.append(("note", f"# CDS sequence length is not a multiple of 3: {file}"))   # :114-119
            q = 0
            for label, hits in zip(labels, fi["hits"]):
                if hits == 0:
                    actions.append(("note", f"# Pop {label} string was not found in fasta headers."))
                else:
                    pending.append((len(actions), fi["locus"], q, file, fi["L"], label, hits))
                    actions.append(None)
                    q += 1
        else:
            actions.append(("raise", api.PolyFastaError(st, "cannot process %s" % path)))
            break
    t2 = time.perf_counter()
    if pending:
        if os.environ.get("POLYFASTA_TIMING"):  # the three steps of Batch.run, timed one by one
            ta = time.perf_counter()
            batch.stage()
            tb = time.perf_counter()
            batch.scan(jc, cds)
            tc = time.perf_counter()
            batch.release()
            print("  gpu run: stage %.1f ms, scan %.1f ms, release %.1f ms" % ((tb - ta) * 1e3, (tc - tb) * 1e3, (time.perf_counter() - tc) * 1e3), file=sys.stderr)
        else:
            batch.run(jc, cds)
        t3 = time.perf_counter()
        for pos, locus, q, file, seqlen, label, n in pending:
            if cds:
                line = cds_line(file, seqlen, label, n, batch.result(locus, q), batch.result_cds(locus, q))
            else:
                line = noncds_line(file, seqlen, label, n, batch.result(locus, q))
            actions[pos] = ("row", line, file, label)
    else:
        t3 = t2
    if os.environ.get("POLYFASTA_TIMING"):
        print("chunk of %d files: add_files %.1f ms (%d threads), plan %.1f ms, gpu run %.1f ms, rows %.1f ms" %
              (len(paths), (t1 - t0) * 1e3, threads, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (time.perf_counter() - t3) * 1e3), file=sys.stderr)
    return [a for a in actions if a is not None]


def emit(actions, sink):
    for a in actions:
        if a[0] == "stdout":
            print(a[1])
        elif a[0] == "note":
            sink.note(a[1])
        elif a[0] == "row":
            sink.row(a[1], a[2], a[3])
        elif a[0] == "raise":
            raise a[1]


def _file_size(p):
    try:
        return os.path.getsize(p)
    except OSError:
        return 0


def devices_from_env():
    """POLYFASTA_DEVICES=0,1,... (or POLYFASTA_DEVICE=0): the GPUs the --dir loci are spread over, round-robin by chunk"""
    spec = os.environ.get("POLYFASTA_DEVICES") or os.environ.get("POLYFASTA_DEVICE") or "0"
    return [int(x) for x in spec.split(",") if x.strip() != ""]


_SLOTS = {}


def run_files(paths, cds, jc, popkeys, sink):
    """the per-file loop of PolyFastA.py:104-140 over sorted paths, in chunks; chunk i runs on device i mod G and the
    rows are emitted in the reference's order"""
    from concurrent.futures import ThreadPoolExecutor
    from . import parallel
    devs = devices_from_env()
    chunks = [ps for _, ps in parallel.plan_units(paths, [_file_size(p) for p in paths], None, BATCH_FILES, BATCH_BYTES)]
    if len(chunks) > 1 and os.environ.get("POLYFASTA_DOUBLE_BUFFER", "1") != "0":
        devs = [d for d in devs for _ in (0, 1)]   # two slots per GPU: host staging of chunk i+1 overlaps the GPU pass of chunk i
    state = _SLOTS   # contexts and pinned staging buffers outlive one call (a process that runs several directories reuses them)

    def work(ci):
        slot = parallel.chunk_owner(ci, len(devs))   # one context + batch per listed device slot (a context is not re-entrant)
        key = (slot, devs[slot])
        if key not in state:
            ctx = api.Context(devs[slot])
            state[key] = (ctx, api.Batch(ctx))
        ctx, batch = state[key]
        threads = max(1, (os.cpu_count() or 1) // len(devs))
        return process_chunk_native(ctx, batch, chunks[ci], jc, popkeys, threads, cds)

    if len(devs) == 1 or len(chunks) == 1:
        for ci in range(len(chunks)):
            emit(work(ci), sink)
    else:
        # one worker thread per device (a context is not re-entrant); ctypes releases the GIL during library calls
        pools = [ThreadPoolExecutor(max_workers=1) for _ in devs]
        futs = [pools[parallel.chunk_owner(ci, len(devs))].submit(work, ci) for ci in range(len(chunks))]
        for f in futs:
            emit(f.result(), sink)
        for pl in pools:
            pl.shutdown()


def run_files_ranks(paths, cds, jc, popkeys, sink, group):
    """the same loop under torchrun (one process per GPU).  Units in the reference's order: a file of at least
    SHARD_MIN_BYTES is ONE alignment scanned by ALL ranks in column shards (exchange fused into the kernels); the other
    files are grouped in chunks and chunk j goes to rank j mod world with no collective.  Rank 0 gathers the
    finished rows and prints them in order."""
    from . import parallel
    ctx = api.Context(group.device)
    batch = api.Batch(ctx)
    threads = max(1, (os.cpu_count() or 1) // group.world)
    ctx.set_host_threads(threads)
    units = parallel.plan_units(paths, [_file_size(p) for p in paths], SHARD_MIN_BYTES, BATCH_FILES, BATCH_BYTES)
    mine, j = [], 0
    for ui, (kind, ps) in enumerate(units):
        if kind == "all":
            fasta = group.parse_once(ps[0], threads)
            acts = process_chunk(ctx, None, [(ps[0], fasta)], cds, jc, popkeys, group=group)
            if group.rank == 0:
                mine.append((ui, acts))
            continue
        if parallel.chunk_owner(j, group.world) == group.rank:
            mine.append((ui, process_chunk_native(ctx, batch, ps, jc, popkeys, threads, cds)))
        j += 1
    group.close(ctx)
    gathered = parallel.gather_rows(mine)
    if group.rank == 0:
        for _, acts in gathered:
            emit(acts, sink)


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.file == "" and args.dir == "." and not args.pipe:
        print("Both --file/-f and --dir/-d were not found.")
        r = input("Do you wish to run polySFS on all files in the current directory? [y|n]: ")
        if r != "y":
            parser.error(parser.print_help())
    if len(args.file) != 0 and args.dir != ".":
        parser.error("Run with either --file/-f or --dir/-d, but not both")
    elif len(args.file) == 0 and args.dir != ".":
        args.file = [args.dir + "/" + i for i in os.listdir(args.dir)]
    elif len(args.file) == 0 and args.dir == "." and args.pipe:
        args.file = [args.name] if args.name else ["stdin"]
    sink = Sink(args.out, args.silent)
    group = RankGroup() if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None   # torchrun: one process per GPU
    if group is not None and group.rank != 0 and args.pipe:
        return 0   # stdin belongs to rank 0
    if group is not None and args.pipe:
        group = None
    if not args.silent and (group is None or group.rank == 0):
        sink.header(args.cds)
    popkeys = args.pops.split(",") if args.pops else None
    paths = sorted(args.file)
    if paths and args.pipe:
        if not (args.file[0] == "stdin" or args.file[0] == args.name):
            parser.error("A FASTA file or multiple files cannot be used with the --pipe argument.")
        data = sys.stdin.buffer.read()
        ctx = api.default_context(devices_from_env()[0])
        batch = api.Batch(ctx)
        for path in paths:   # the reference reads stdin once per listed name; the second read is empty
            try:
                fasta = api.Fasta.from_bytes(data)
            except api.NotFasta as e:
                fasta = e
            emit(process_chunk(ctx, batch, [(path, fasta)], args.cds, args.jc, popkeys), sink)
            data = b""
    elif paths and group is not None:
        run_files_ranks(paths, args.cds, args.jc, popkeys, sink, group)
    elif paths:
        run_files(paths, args.cds, args.jc, popkeys, sink)
    if len(args.out) != 0 and not args.silent and (group is None or group.rank == 0):
        print("")
    return 0


if __name__ == "__main__":
    sys.exit(main())
