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


def process_alignment(ctx, fasta, file, cds, jc, popkeys, sink):
    """one file: equal-length check, %3 warning, population split, GPU scans, rows (PolyFastA.py:109-140)"""
    if fasta.seqlen < 0:
        sink.note(f"# Sequences do not have the same length: {file}")
        return
    seqlen = fasta.seqlen
    if cds and seqlen % 3 != 0:
        sink.note(f"# CDS sequence length is not a multiple of 3: {file}")
    if popkeys is None:
        plan = [("NA", list(range(fasta.nseq)))]
    else:
        heads = fasta.headers
        plan = [(key, [i for i, h in enumerate(heads) if key in h]) for key in popkeys]
    found = [(label, rows) for label, rows in plan if rows]
    lines = []
    if found:
        aln = api.Alignment.from_fasta(ctx, fasta)
        try:
            aln.set_pops([rows for _, rows in found])
            site = aln.site_stats()
            cdsst = aln.cds_stats() if cds else None
        finally:
            aln.free()
        lines = format_rows(ctx, file, seqlen, cds, jc, [(label, len(rows)) for label, rows in found], site, cdsst)
    it = iter(lines)
    for label, rows in plan:
        if not rows:
            sink.note(f"# Pop {label} string was not found in fasta headers.")
        else:
            sink.row(next(it), file, label)


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
    if not args.silent:
        sink.header(args.cds)
    popkeys = args.pops.split(",") if args.pops else None
    ctx = None
    for path in sorted(args.file):
        if len(args.file) > 0 and args.pipe and not (args.file[0] == "stdin" or args.file[0] == args.name):
            parser.error("A FASTA file or multiple files cannot be used with the --pipe argument.")
        try:
            if args.pipe:
                fasta = api.Fasta.from_bytes(sys.stdin.buffer.read())
            else:
                fasta = api.Fasta.from_file(path)
        except api.NotFasta:
            print(f"# file {path} is not FASTA!")
            continue
        if ctx is None:
            ctx = api.default_context(int(os.environ.get("POLYFASTA_DEVICE", "0")))
        try:
            process_alignment(ctx, fasta, path.split("/")[-1], args.cds, args.jc, popkeys, sink)
        finally:
            fasta.close()
    if len(args.out) != 0 and not args.silent:
        print("")
    return 0


if __name__ == "__main__":
    sys.exit(main())
