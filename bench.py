#!/usr/bin/env python
"""Benchmark of the seed-scanning hot path (BASELINE.json metric: Gbp/s scanned, motifs 2-100).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU scan on the host cores

A step = one pass of the hot path (pack -> match planes -> perfect / substitution / anchored seed machines -> ordered
candidate streams) over one batch of synthetic input. Workload at every N: each GPU scans its own chr21-scale
synthetic contig (BASELINE.json configs[1], 46.7 Mbp, seed 21 + rank) -> weak scaling, no collective on the data path.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_LO, M_HI = 2, 100
N_SHIFTS, N_MOTIFS = 102, 99
OPS_PER_BASE = (12 * N_SHIFTS + 39 * N_MOTIFS) / 32.0   # algorithmic 32-lane word-ops per base
PACKED_BYTES_PER_BASE = 0.375                           # three 1-bit planes
ASCII_BYTES_PER_BASE = 1.375                            # ASCII in, planes out
DEFAULT_BASES = 46_700_000


def workload_name(bases):
    return "C2 chr21-scale synthetic contig, %.1f Mbp per GPU, random background + planted perfect/impure repeats + 2 N runs, -m 2 -M 100" % (bases / 1e6)


def make_contig(bases, rank):
    from ribbit_b200 import synth
    return synth.contig_c2(bases, seed=21 + rank)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own scan (oracle/_ref, unmodified reference sources) or, failing that, the oracle port
# ------------------------------------------------------------------------------------------------------------------
def _ref_scan_seconds(seq, workdir, tag):
    """Wall seconds of the reference's scan stage (pack -> shift-XOR -> perfect -> substitution -> anchors -> anchored,
    fasta_utils.cpp:78-170) on `seq`: oracle/_ref/ribbit_ref_cp with RB_CP_STOP_AFTER_CP2 returns before the per-seed
    stage; without RB_CP_OUT it logs nothing."""
    from ribbit_b200 import synth
    fa = os.path.join(workdir, "s%s.fa" % tag)
    synth.write_fasta(fa, [seq])
    exe = os.path.join(ROOT, "oracle", "_ref", "ribbit_ref_cp")
    env = dict(os.environ, RB_CP_STOP_AFTER_CP2="1")
    env.pop("RB_CP_OUT", None)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "-i", fa, "-o", os.path.join(workdir, "o%s.bed" % tag), "-m", str(M_LO), "-M", str(M_HI)],
                       env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference exited with %d" % r.returncode)
    return dt


def _port_scan_seconds(seq):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_util as ou
    t0 = time.perf_counter()
    ou.scan_count(seq, M_LO, M_HI)
    return time.perf_counter() - t0


def cpu_scan_throughput(contig, sample_bases, nproc):
    """Scans `nproc` disjoint slices of `sample_bases` bases of the workload concurrently, one single-threaded
    reference process per slice (the reference has no threads). Returns (Gbp/s, kind, cores, description)."""
    have_ref = os.access(os.path.join(ROOT, "oracle", "_ref", "ribbit_ref_cp"), os.X_OK)
    L = len(contig)
    # slices from the repeat-bearing part of the contig (the first 50 kb are an N run)
    starts = [min(max(0, L - sample_bases), 100_000 + i * sample_bases) for i in range(nproc)]
    slices = [contig[s:s + sample_bases] for s in starts]
    times = [0.0] * nproc
    errs = []
    with tempfile.TemporaryDirectory() as td:
        def work(i):
            try:
                times[i] = _ref_scan_seconds(slices[i], td, str(i)) if have_ref else _port_scan_seconds(slices[i])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        t0 = time.perf_counter()
        if have_ref:
            th = [threading.Thread(target=work, args=(i,)) for i in range(nproc)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        else:
            nproc = 1
            work(0)
        wall = time.perf_counter() - t0
    if errs:
        raise errs[0]
    total = sum(len(s) for s in slices[:nproc])
    kind = "reference" if have_ref else "port"
    what = ("%d x %.2f Mbp slices of the workload contig, one single-threaded %s process per slice, scan stage only "
            "(pack..anchored seeds), wall %.1f s" % (nproc, sample_bases / 1e6,
                                                     "reference (oracle/_ref/ribbit_ref_cp, stop after the scan)" if have_ref
                                                     else "oracle port", wall))
    return total / wall / 1e9, kind, nproc, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncpu = os.cpu_count() or 1
    contig = make_contig(min(args.bases, 12_000_000), 0)
    vals = []
    what = ""
    for step in range(args.warmup + args.steps):
        # bounded sample per step: each process scans 0.5 Mbp (~3 s of single-core work)
        v, kind, cores, what = cpu_scan_throughput(contig, args.ref_sample, ncpu)
        if step >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "Gbp/s scanned (motifs 2-100)", "value": value, "unit": "Gbp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (cores * args.ref_sample / 1e9) / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.bases), "sample": what},
        "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ribbit_b200 import scan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the scan has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when the first communicator is created; this program's stdout
        # carries exactly one JSON line, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    contig = make_contig(args.bases, rank)
    L = len(contig)
    host = torch.empty(L + 64, dtype=torch.uint8).pin_memory()
    host[:L] = torch.frombuffer(bytearray(contig), dtype=torch.uint8)
    host_np = host.numpy()
    dev = host.cuda(non_blocking=False)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    sc = scan.Scanner(M_LO, M_HI, device=local)
    int_peak = sc.int_peak()

    # ---- device-resident throughput: the input is in HBM, the three ordered streams stay in HBM ------------------
    sc.load_device(dev.data_ptr(), [L], keepalive=dev)
    sampler = ClockSampler(local)
    sampler.start()                       # clocks under load: sampled from the warm-up steps to the end of the timed region
    for _ in range(args.warmup):
        sc.scan_device()
    counts = sc.counts()
    barrier()
    per_step, scan_ms, pack_ms, merge_ms, launches, restarts = [], [], [], [], 0, 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                     # L2 flush between timed iterations
        torch.cuda.synchronize()
        sc.scan_device()                  # CUDA events on the library's stream bracket the kernels
        t = sc.timing()
        per_step.append(t["total_ms"]); scan_ms.append(t["scan_ms"]); pack_ms.append(t["pack_ms"]); merge_ms.append(t["merge_ms"])
        launches += t["launches"]
        restarts += t["restarts"]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    dev_ms = float(sum(per_step))
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tmax.item())
    value = world * L * args.steps / (dev_ms_max * 1e-3) / 1e9

    # ---- end to end through the public API with host buffers: every step copies its ASCII from pinned host memory
    # (rb_load_contigs), runs the kernels and copies its three streams back (rb_scan). Steps go through
    # ribbit_b200.pipeline.ScanPipeline: three contexts on the GPU, so the copies of one step overlap the kernels of the
    # next, as when a genome is scanned contig by contig. All K results are complete inside the timed region.
    from ribbit_b200 import pipeline
    pipe = pipeline.ScanPipeline(M_LO, M_HI, device=local, depth=3, compact=True)
    for f in [pipe.submit_flat(host_np[:L + 1], [L]) for _ in range(4)]:
        f.result()
    barrier()
    t0 = time.perf_counter()
    futs = [pipe.submit_flat(host_np[:L + 1], [L]) for _ in range(args.steps)]
    res = None
    for f in futs:
        res = f.result()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * L * args.steps / float(te.item()) / 1e9
    d2h = int(sum(len(res[s][0]) for s in range(3)) * 8 + 3 * 2 * 8)
    # the same without overlap: one context, load -> scan -> fetch back to back
    sc2 = scan.Scanner(M_LO, M_HI, device=local)
    sc2.load_flat(host_np[:L + 1], [L]); sc2.scan(copy=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sc2.load_flat(host_np[:L + 1], [L])
        sc2.scan(copy=False)
    e2e_serial = L * args.steps / (time.perf_counter() - t0) / 1e9

    # ---- the rows next to the path that are built, measured outside the timed regions, rank 0 only:
    # K0 rb_load_fasta (80-column FASTA text of the same contig, pinned host memory -> contigs on the device) and
    # K7 rb_motif_rows (seeds = the scan's own kept anchored candidates with motif sizes > 10, N-truncated by K5)
    next_rows = None
    if rank == 0:
        body = np.frombuffer(contig, dtype=np.uint8)
        full = (L // 80) * 80
        lines = np.concatenate([body[:full].reshape(-1, 80), np.full((full // 80, 1), 10, np.uint8)], axis=1).reshape(-1)
        text_np = np.concatenate([np.frombuffer(b">chr21 synthetic\n", np.uint8), lines, body[full:], np.frombuffer(b"\n", np.uint8)])
        text = torch.empty(len(text_np), dtype=torch.uint8).pin_memory()
        text.numpy()[:] = text_np
        t_fa = []
        for _ in range(4):
            t0 = time.perf_counter()
            names, lens = sc2.load_fasta(text.numpy())
            t_fa.append(time.perf_counter() - t0)
        assert names == ["chr21"] and lens.tolist() == [L]
        sc2.scan_device()
        assert sc2.counts() == counts, "streams differ after rb_load_fasta"
        a = sc2.fetch(copy=False)[2][0]
        a = a[(a["mlen"] > 10) & (a["flags"] == 0)][:400_000]
        seeds = np.stack([np.zeros(len(a), np.int32), a["start"], a["end"], a["mlen"].astype(np.int32)], axis=1).astype(np.int32)
        info = sc2.filter_seeds(seeds)
        seeds[:, 2] = np.minimum(seeds[:, 1] + info[:, 0], L)
        t_mr = []
        for _ in range(3):
            t0 = time.perf_counter()
            rows = sc2.motif_rows(seeds)
            t_mr.append(time.perf_counter() - t0)
        next_rows = {"k0_load_fasta": {"text_bytes": int(len(text_np)), "ms": min(t_fa[1:]) * 1e3, "gbp_per_s": L / min(t_fa[1:]) / 1e9,
                                       "what": "rb_load_fasta from pinned host memory: H2D of the text + 3 kernels + header table D2H"},
                     "k7_motif_rows": {"seeds": int(len(seeds)), "ms": min(t_mr) * 1e3, "seeds_per_s": len(seeds) / min(t_mr),
                                       "scored": int((rows[:, 1] > 0).sum()),
                                       "what": "rb_motif_rows on the scan's kept anchored candidates with mlen > 10 (whole call: H2D of the seeds, kernel, D2H)"}}

    if rank == 0:
        scan_ms_avg = float(np.mean(scan_ms))
        ops_per_launch = OPS_PER_BASE * L                   # lane-operations (one 32-bit op in one lane = 32 bases x 1 op)
        achieved = ops_per_launch / (scan_ms_avg * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_achieved = PACKED_BYTES_PER_BASE * L / (scan_ms_avg * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json"))).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        ncpu = os.cpu_count() or 1
        cpu_v, cpu_kind, cpu_cores, cpu_what = cpu_scan_throughput(contig, args.cpu_sample, min(ncpu, 8))
        line = {
            "metric": "Gbp/s scanned (motifs 2-100)", "value": value, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": workload_name(L), "bases_per_gpu": L, "min_mlen": M_LO, "max_mlen": M_HI,
                       "l2": "flushed (512 MiB write) between timed steps", "timing": "CUDA events on the library stream, max over ranks",
                       "candidates_per_step": counts, "stage_ms": {"pack": float(np.mean(pack_ms)), "scan": scan_ms_avg,
                                                                   "merge": float(np.mean(merge_ms))},
                       "wall_ms_bracket": wall_ms, "warmup_restarts_per_step": restarts / args.steps},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gbp/s", "h2d_bytes_per_step": int(L), "d2h_bytes_per_step": d2h,
                    "what": "per step: rb_load_contigs (pinned host ASCII -> HBM) + rb_scan_device (kernels) + rb_fetch_compact (D2H of the three streams as 8-byte records); steps pipelined over 3 contexts (ribbit_b200.pipeline)",
                    "serial_one_context_gbps_per_gpu": e2e_serial},
            "gpu_launches": launches,
            "roofline": {"bound": "int", "kernel": "scan_kernel<32>", "achieved": achieved / 1e12, "peak": int_peak / 1e12,
                         "unit": "Tlaneop/s", "frac": achieved / int_peak,
                         "note": "algorithmic (12*NSHIFTS+39*NMOTIFS)/32 = %.1f word-ops per base; peak = LOP3+SHF microbenchmark measured in this run; the HBM bound is ~100x looser" % OPS_PER_BASE,
                         "traffic": traffic,
                         "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_achieved / hbm_peak, "peak_source": "measured" if peaks else "fallback"}},
            "cpu_baseline": {"value": cpu_v, "unit": "Gbp/s", "cores": cpu_cores, "kind": cpu_kind, "sample": cpu_what},
            "next_rows": next_rows,
        }
        print(json.dumps(line), flush=True)
    sc.close(); sc2.close(); pipe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--bases", type=int, default=DEFAULT_BASES, help="bases per GPU")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000, help="bases per CPU-baseline process")
    ap.add_argument("--ref-sample", type=int, default=500_000, help="bases per process and step for --impl reference")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()
