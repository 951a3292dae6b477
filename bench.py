#!/usr/bin/env python3
"""Benchmark of the PolyFastA hot path on B200:  python bench.py --gpus N --steps K --warmup W  [--impl reference]

Workload (BASELINE.json configs[3], the one the metric is quoted on): synthetic non-coding alignment of 10,000
sequences x 10 Mb (1e11 aligned bases), pure ACGT, ~5 % segregating sites with a 1/k spectrum, 1 % of them
tri-allelic; generated directly in packed form on the device (polyfasta_b200.synth is the numpy twin).  With N GPUs the
COLUMNS are split in N contiguous ranges (strong scaling: the total stays 1e11 bases) and the int64 vector [S, H, SFS] is
summed with ONE NCCL all-reduce per step.

  step        one pass of the site scan (K2) over the resident shard + the all-reduce (N > 1)
  value       aligned bases/s = n*L*K / (device time of the K steps, max over ranks); inputs resident in HBM
  roofline    the site-scan kernel: algorithmic bytes (what the scan must read: two bit-planes = 0.25 B/base for a pure-ACGT
              shard, three = 0.375 B/base otherwise) / its mean CUDA-event duration inside the timed region
  e2e         the same metric from HOST text: per step, the pinned row-major text of a column slice -> H2D -> K1 encode ->
              K2 scan -> all-reduce -> D2H of the vector -> K5 finalise (fp64 on device) -> Python tuple
  cpu_baseline / --impl reference: the CPU oracle port (oracle/c, C + OpenMP, all host threads) on a bounded sample of the
              same workload.  The reference itself is pure Python and cannot travel to the GPU box; its own speed measured
              in the build container is ~1e7 bases/s on one core.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SEQ = 10_000
N_SITES = 10_000_000
SEED = 4
P_SEG_PPM = 50_000
TRI_PPM = 10_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=N_SEQ)
    ap.add_argument("--sites", type=int, default=N_SITES)
    ap.add_argument("--e2e-sites", type=int, default=1_000_000, help="column slice used for the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample-sites", type=int, default=100_000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"],
                    help="N > 1: 'fused' = the scan kernel's last block sums the shard vectors over NVLink peer memory (pfa_xchg, the product "
                         "path); 'nccl' = scan kernel + torch.distributed all_reduce (comparison only)")
    ap.add_argument("--force-validity", action="store_true", help="make the scan read the validity plane too (0.375 B/base)")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every few ms DURING the timed region (B200_PROFILING.md)"""

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.sm_max, self.power = [], 0, None, []
        self.stop_flag = threading.Event()
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def summary(self):
        names = []
        if self.h is not None:
            nv = self.nv
            for nm, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                            ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if self.reasons & bit:
                    names.append(nm)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max, "reasons": names,
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


def host_threads():
    """all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not cap the CPU arm)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_rate(n, sample_sites, seconds, min_passes=2):
    """the oracle port (C + OpenMP, all host threads) on columns [0, sample_sites) of the workload -> bases/s"""
    from oracle import c_oracle as co
    co.build()
    threads = host_threads()
    mat = co.synth_text(SEED, n, N_SITES, P_SEG_PPM, TRI_PPM, 0, sample_sites, threads=threads)
    co.site_stats(mat, threads=threads)  # warm-up
    t0 = time.perf_counter()
    passes = 0
    while passes < min_passes or time.perf_counter() - t0 < seconds:
        r = co.site_stats(mat, threads=threads)
        passes += 1
    dt = time.perf_counter() - t0
    return n * sample_sites * passes / dt, threads, passes, r


def run_reference(args, rank):
    """--impl reference: the CPU implementation of the path on the box's host cores; rank 0 only"""
    if rank != 0:
        return
    from oracle import c_oracle as co
    co.build()
    threads = host_threads()
    n, sample = args.n, args.cpu_sample_sites
    mat = co.synth_text(SEED, n, args.sites, P_SEG_PPM, TRI_PPM, 0, sample, threads=threads)
    for _ in range(args.warmup):
        co.site_stats(mat, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.site_stats(mat, threads=threads)
    dt = time.perf_counter() - t0
    value = n * sample * args.steps / dt
    sample_txt = "columns [0,%d) of the %d x %d alignment per step (%.1e bases), C+OpenMP oracle port" % (sample, n, args.sites, n * sample)
    line = {
        "impl": "reference", "metric": "aligned bases/sec (seqs x sites)", "value": value, "unit": "bases/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "C4: synthetic non-coding alignment %d seqs x %d sites, site scan (S, H, folded SFS)" % (n, args.sites),
                   "sample": sample_txt},
        "cpu_baseline": {"value": value, "unit": "bases/s", "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python (not on this box); its own functions ran at ~1e7 bases/s on one core in the build container (BASELINE.md)",
    }
    print_json(line)


def main():
    args = parse_args()
    # rank 0 prints exactly ONE line on stdout; anything a library writes to fd 1 (NCCL's version banner) goes to stderr
    global print_json
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def print_json(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import polyfasta_b200 as pf
    from polyfasta_b200 import api, parallel, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, L = args.n, args.sites
    ctx = pf.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    def shard(total, r):
        return parallel.shard_columns(total, world, r)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.cuda.stream(stream):
        c0, c1 = shard(L, rank)
        aln = pf.Alignment.synthetic(ctx, n, L, SEED, P_SEG_PPM, TRI_PPM, c0, c1)
        if args.force_validity:
            aln.force_validity(True)
        planes_read = 3 if (aln.has_invalid or args.force_validity) else 2
        out = torch.zeros(aln.site_len(), dtype=torch.int64, device="cuda")
        fused = world > 1 and args.collective == "fused"
        xchg = parallel.connect_exchange(ctx, aln.site_len()) if fused else None
        ctx.sync()

        def step(ev=None):
            if ev:
                ev[0].record(stream)
            if fused:
                aln.site_stats_xchg(xchg, out.data_ptr())     # ONE launch: K2 + sum over the shards of all ranks
            else:
                aln.site_stats_device(out.data_ptr())
            if ev:
                ev[1].record(stream)
            if world > 1 and not fused:
                dist.all_reduce(out)

        for _ in range(args.warmup):
            step()
        barrier()
        launches0 = ctx.launch_count
        sampler = ClockSampler(local_rank)
        sampler.start()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            # device-side rendezvous on the timing stream: the ranks leave the host barrier up to a millisecond apart, and
            # without it the first timed step of the early ranks would just measure that host skew
            sync_word = torch.zeros(1, dtype=torch.int64, device="cuda")
            if fused:
                xchg.allreduce(sync_word.data_ptr(), 1)
            else:
                dist.all_reduce(sync_word)
        t_start.record(stream)
        for i in range(args.steps):
            step(kev[i])
        t_end.record(stream)
        barrier()
        sampler.stop_flag.set()
        sampler.join()
        launches = ctx.launch_count - launches0
        elapsed_ms = max_over_ranks(t_start.elapsed_time(t_end))
        kernel_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in kev) / args.steps)
        result = out.cpu().numpy().copy()
        if fused and xchg.timed_out():
            raise SystemExit("the NVLink exchange timed out on rank %d (a rank did not arrive)" % rank)
        if os.environ.get("PFA_XCHG_STAMPS"):
            per = [a.elapsed_time(b) for a, b in kev]
            gaps = [kev[i][0].elapsed_time(kev[i + 1][0]) for i in range(args.steps - 1)]
            sys.stderr.write("rank %d kernel ms per step %s\nrank %d step-to-step ms %s\n" %
                             (rank, ["%.3f" % x for x in per], rank, ["%.3f" % x for x in gaps]))
            if fused:
                sys.stderr.write("rank %d exchange stages (ns): %s\n" % (rank, xchg.stamps()))

    value = n * L * args.steps / (elapsed_ms * 1e-3)
    my_sites = c1 - c0
    algo_bytes = n * my_sites * planes_read / 8.0
    peak, peak_kind = measured_peak()
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get("C4 n=%d sites=%d gpus=%d planes=%d" % (n, L, world, planes_read))
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass

    # ---- parity at full size: the generator's closed form (numpy) ----
    parity = "skipped"
    if rank == 0 and not args.no_check:
        want = synth.expected_site_stats(SEED, n, L, P_SEG_PPM, TRI_PPM)
        got = (int(result[0]), int(result[1]), [int(x) for x in result[2:]])
        if got != (want["S"], want["H"], want["sfs"]):
            raise SystemExit("PARITY FAILURE at full size: got S=%d H=%d, closed form S=%d H=%d" % (got[0], got[1], want["S"], want["H"]))
        parity = "S, H and the folded SFS of the %d x %d alignment equal the generator's closed form (S=%d)" % (n, L, want["S"])
    fin = ctx.finalize([(n, int(result[0]), int(result[1]), L, True)])[0]
    # the read-only ceiling for exactly these bytes on this GPU (outside the timed region)
    probe_ms = aln.read_probe(planes_read, 5)
    read_only_gbs = aln.packed_bytes / 3 * planes_read / (probe_ms * 1e-3) / 1e9
    aln.free()
    ctx.trim()

    # ---- end to end from host text ----
    e2e = None
    if not args.no_e2e:
        with torch.cuda.stream(stream):
            es = min(args.e2e_sites, L)
            e0, e1 = shard(es, rank)
            cols = e1 - e0
            ld = (cols + 255) // 256 * 256
            d_text = torch.empty((n, ld), dtype=torch.uint8, device="cuda")
            api.synth_text_device(ctx, d_text.data_ptr(), ld, n, SEED, P_SEG_PPM, TRI_PPM, e0, e1)
            ctx.sync()
            h_text = torch.empty((n, ld), dtype=torch.uint8, pin_memory=True)
            h_text.copy_(d_text)
            torch.cuda.synchronize()
            del d_text
            torch.cuda.empty_cache()
            h_out = torch.empty(out.numel(), dtype=torch.int64, pin_memory=True)

            ctx.set_host_threads(max(1, host_threads() // world))   # the ranks of one box share its cores

            def e2e_step():
                a = pf.Alignment.from_host_ptr(ctx, h_text.data_ptr(), n, cols, ld)   # H2D (pinned, chunked) + K1
                if fused:
                    a.site_stats_xchg(xchg, out.data_ptr())                          # K2 + sum over shards (NVLink)
                else:
                    a.site_stats_device(out.data_ptr())                              # K2
                    if world > 1:
                        dist.all_reduce(out)
                h_out.copy_(out, non_blocking=True)                                  # D2H
                stream.synchronize()
                r = ctx.finalize([(n, int(h_out[0]), int(h_out[1]), es, True)])[0]   # K5
                a.free()
                return r

            e2e_step()
            barrier()
            l0 = ctx.launch_count
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                r = e2e_step()
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0)
            e2e_launches = (ctx.launch_count - l0) // args.e2e_steps
            if rank == 0 and not args.no_check:
                want = synth.expected_site_stats(SEED, n, L, P_SEG_PPM, TRI_PPM, 0, es)
                if (int(h_out[0]), int(h_out[1])) != (want["S"], want["H"]):
                    raise SystemExit("PARITY FAILURE in the end-to-end leg")
            ing = ctx.ingest_stats()   # the last step's upload on this rank (every rank's shard has the same shape)
            e2e = {"value": n * es * args.e2e_steps / dt, "unit": "bases/s",
                   "h2d_bytes_per_step": (ing["h2d_text_bytes"] + ing["h2d_packed_bytes"]) * world,
                   "d2h_bytes_per_step": (out.numel() * 8 + 32) * world, "ms_per_step": dt / args.e2e_steps * 1e3,
                   "sample": "columns [0,%d) of the workload as row-major text in pinned host memory (%.1e bases/step); "
                             "hybrid ingest (column chunks either as text over PCIe -> K1, or packed 4 bases/byte by host threads -> packed "
                             "K1) + K2 + sum over shards + D2H + K5 inside the timed region" % (es, n * es),
                   "gpu_launches_per_step": e2e_launches, "ingest": ing, "host_text_bytes_per_step": n * cols * world, "result": [r[0], r[1], r[2], r[3]]}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, threads, passes, r = cpu_oracle_rate(n, args.cpu_sample_sites, args.cpu_seconds)
        cpu = {"value": rate, "unit": "bases/s", "cores": threads, "kind": "port",
               "sample": "columns [0,%d) of the workload (%.1e bases) x %d passes, C+OpenMP oracle port (oracle/c)" %
                         (args.cpu_sample_sites, n * args.cpu_sample_sites, passes)}

    if rank == 0:
        line = {
            "metric": "aligned bases/sec (seqs x sites)", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "C4: synthetic non-coding alignment %d seqs x %d sites, site scan (S, H, folded SFS)" % (n, L),
                       "parallelism": ("columns split in %d contiguous ranges; " % world + ("the scan kernel's last block sums the int64 shard vectors "
                                       "over NVLink peer memory (fused, no collective call)" if fused else "1 NCCL int64 all-reduce per step"))
                       if world > 1 else "1 GPU",
                       "l2": "inputs larger than L2 (%.1f GB of planes read per GPU per step)" % (algo_bytes / 1e9),
                       "planes_read": planes_read, "seed": SEED, "p_seg_ppm": P_SEG_PPM, "tri_ppm": TRI_PPM},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback",
                         "kernel": "pfa_site_scan_tma_kernel<16,5,HAS_V,512>" if n == N_SEQ else "pfa_site_scan_*", "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": algo_bytes,
                         "read_only_peak": read_only_gbs, "frac_of_read_only_peak": achieved / read_only_gbs,
                         "read_only_peak_kind": "pfa_read_probe_kernel over the same planes (padded bytes), measured in this run"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(), "parity": parity,
            "result": {"S": int(result[0]), "H": int(result[1]), "pi_site_jc": fin[1], "theta_site": fin[2], "tajimasD": fin[3]},
        }
        print_json(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
