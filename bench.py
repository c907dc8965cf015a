#!/usr/bin/env python
"""Benchmark of the seed-scanning hot path (BASELINE.json metric: Gbp/s scanned, motifs 2-100, at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU scan on the host cores

Workload: C3, BASELINE.json configs[2] — the human-genome-scale synthetic the metric is quoted on: 24 contigs with
hg38-like lengths, 3.1 Gbp, -m 2 -M 100 (ribbit_b200/workloads.py). STRONG scaling: the genome is fixed, its contigs
are spread over the ranks by (contig, chunk) (ribbit_b200/shard.py plan_parts: whole contigs longest-first, then the tail
of the largest contigs handed over as 32-base word ranges, rb_set_word_range); no collective on the data path.
A step = one pass of the hot path (pack -> match planes -> perfect / substitution / anchored seed machines -> ordered
candidate streams) over the rank's share of the genome.

Parity gate inside the run: the kept candidates of a 1 Mbp window of every contig and of the whole chr21-size contig are
digested (md5, call order) and compared with the digests of the unmodified reference (tests/golden/c3_digests.json,
made by tests/golden/make_golden_c3.py with oracle/_ref/ribbit_ref_cp); a mismatch exits non-zero, no number is printed.
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
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_LO, M_HI = 2, 100
PACKED_BYTES_PER_BASE = 0.375                           # three 1-bit planes


def ops_per_base(m_lo, m_hi):
    """algorithmic 32-lane word-ops per base = (12 NSHIFTS + 39 NMOTIFS) / 32."""
    s_lo = m_lo - 2 if m_lo > 2 else 1
    return (12 * (m_hi + 2 - s_lo + 1) + 39 * (m_hi - m_lo + 1)) / 32.0


def workload_name(scale):
    return ("C3 human-genome-scale synthetic: 24 contigs with hg38-like lengths, %.2f Gbp, contig i = C2 generator with seed 100+i "
            "(random background, 300 planted perfect/impure repeats per Mbp, 2 N runs, 5%% soft-masked), -m 2 -M 100" %
            (sum(_wl().c3_lengths(scale)) / 1e9))


def _wl():
    from ribbit_b200 import workloads
    return workloads


def _gen(task):
    """Worker of the generator pool (runs before any CUDA call of the parent)."""
    from ribbit_b200 import synth, workloads
    kind = task[0]
    if kind == "c3":
        return workloads.c3_contig(task[1], task[2])
    if kind == "c2":
        return synth.contig_c2(task[1], seed=21)
    if kind == "c5":
        return b"".join(synth.contigs_c5(n=task[1], length=1000, seed=5))
    raise ValueError(kind)


def generate(tasks, nproc):
    if not tasks:
        return []
    with ProcessPoolExecutor(max(1, min(nproc, len(tasks)))) as ex:
        return list(ex.map(_gen, tasks))


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


def cpu_scan_throughput(contig, sample_bases, nproc, what_contig):
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
    what = ("%d x %.2f Mbp slices of %s, one single-threaded %s process per slice, scan stage only "
            "(pack..anchored seeds), wall %.1f s" % (nproc, sample_bases / 1e6, what_contig,
                                                     "reference (oracle/_ref/ribbit_ref_cp, stop after the scan)" if have_ref
                                                     else "oracle port", wall))
    return total / wall / 1e9, kind, nproc, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = _wl()
    ncpu = os.cpu_count() or 1
    ci = wl.C3_FULL_CONTIG
    contig = wl.c3_contig(ci, min(args.scale, 12e6 / wl.c3_lengths()[ci]))  # 12 Mbp of it are enough to cut the samples from
    vals = []
    what = ""
    for step in range(args.warmup + args.steps):
        # bounded sample per step: each process scans 0.5 Mbp (~3 s of single-core work)
        v, kind, cores, what = cpu_scan_throughput(contig, args.ref_sample, ncpu, "C3 contig %d (first 12 Mbp)" % ci)
        if step >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "Gbp/s scanned (motifs 2-100)", "value": value, "unit": "Gbp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (cores * args.ref_sample / 1e9) / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.scale), "sample": what},
        "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
def kept_rows(res, j, lo=None, hi=None):
    """Rows (start, end, mlen) of the kept records (neither DROPPED nor PSEUDO) of contig j of a compact result
    (Scanner.fetch_compact), per stream; with a window only those wl.GATE_MARGIN bases inside [lo, hi)."""
    wl = _wl()
    out = []
    for s in range(3):
        rec, off, lg = res[s]
        a, b = int(off[j]), int(off[j + 1])
        r = rec[a:b]
        keep = ((r["mf"] >> 12) & 3) == 0
        if lo is not None:
            keep &= r["start"] >= lo + wl.GATE_MARGIN
            keep &= r["start"] <= hi  # cheap pre-selection before the 64-bit arithmetic
        idx = np.flatnonzero(keep)
        start = r["start"][idx].astype(np.int64)
        end = start + r["len"][idx].astype(np.int64)
        if len(lg) and len(idx):
            sel = lg[(lg["index"] >= a) & (lg["index"] < b)]   # candidates of >= 65535 positions: true end in the side list
            if len(sel):
                rel = sel["index"] - a
                pos = np.minimum(np.searchsorted(idx, rel), len(idx) - 1)
                ok = idx[pos] == rel
                end[pos[ok]] = sel["end"][ok]
        mlen = (r["mf"][idx] & 0xFFF).astype(np.int64)
        if lo is not None:
            k2 = end <= hi - wl.GATE_MARGIN
            start, end, mlen = start[k2], end[k2], mlen[k2]
        out.append(np.stack([start, end, mlen], axis=1))
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ribbit_b200 import pipeline, scan, shard
    wl = _wl()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = os.cpu_count() or 1

    # ---- the workload: this rank's parts of the genome, generated before the first CUDA call (the pool forks) ----------
    lengths = wl.c3_lengths(args.scale)
    n_words = [(L + 31) // 32 for L in lengths]
    plan = shard.plan_parts(lengths, world)
    mine = plan[rank]
    need = sorted({c for c, _, _ in mine})
    tasks = [("c3", c, args.scale) for c in need]
    extras = rank == 0 and not args.no_extras
    if extras:
        tasks += [("c2", 46_700_000), ("c5", args.c5_contigs)]
    t_gen0 = time.perf_counter()
    blobs = generate(tasks, max(1, ncpu // world))
    gen_s = time.perf_counter() - t_gen0
    seqs = dict(zip(need, blobs[:len(need)]))
    c2_seq, c5_flat = (blobs[len(need)], blobs[len(need) + 1]) if extras else (None, None)

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

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # pinned host copy of this rank's contigs (offsets 64-byte aligned) and its resident copy in HBM
    offs, total = {}, 0
    for c in need:
        offs[c] = total
        total += (len(seqs[c]) + 1 + 63) // 64 * 64
    host = torch.empty(max(total, 64), dtype=torch.uint8, pin_memory=True)
    host_np = host.numpy()
    for c in need:
        host_np[offs[c]:offs[c] + len(seqs[c])] = np.frombuffer(seqs[c], dtype=np.uint8)
    dev = host.cuda()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    whole = [c for c, a, b in mine if a == 0 and b == n_words[c]]
    ranges = [(c, a, b) for c, a, b in mine if not (a == 0 and b == n_words[c])]
    my_bases = sum(min(32 * b, lengths[c]) - 32 * a for c, a, b in mine)

    # ---- device-resident throughput: the input is in HBM, the three ordered streams stay in HBM ------------------
    # one context scans the rank's whole contigs as one batch, one more per word range of a contig shared with other ranks
    ctxs = []
    if whole:
        sc = scan.Scanner(M_LO, M_HI, device=local)
        sc.load_device(dev.data_ptr(), [lengths[c] for c in whole], offsets=[offs[c] for c in whole], keepalive=dev)
        ctxs.append((sc, whole, None))
    for c, a, b in ranges:
        sc = scan.Scanner(M_LO, M_HI, device=local)
        sc.load_device(dev.data_ptr() + offs[c], [lengths[c]], keepalive=dev)
        sc.set_word_range(a, b)
        ctxs.append((sc, [c], (a, b)))
    int_peak = ctxs[0][0].int_peak() if ctxs else 0.0
    int_peak_modes = ctxs[0][0].int_peak_modes() if ctxs and rank == 0 else None

    sampler = ClockSampler(local)
    sampler.start()                       # clocks under load: sampled from the warm-up steps to the end of the timed region
    for _ in range(args.warmup):
        for sc, _, _ in ctxs:
            sc.scan_device()
    counts = [sum(sc.counts()[s] for sc, _, _ in ctxs) for s in range(3)]

    # ---- parity gate on what the timed steps compute: kept candidates vs the reference's digests -------------------
    gate_rows = {}   # (contig, part first word) -> {"win": [rows per stream], "full": [...] or None}
    if args.scale == 1.0 and not args.no_gate:
        for sc, cs, rng in ctxs:
            res = sc.fetch_compact(copy=False)
            for j, c in enumerate(cs):
                lo, hi = wl.c3_window(c, lengths[c])
                gate_rows[(c, rng[0] if rng else 0)] = {"win": kept_rows(res, j, lo, hi),
                                                        "full": kept_rows(res, j) if c == wl.C3_FULL_CONTIG else None}
            del res
    parity = "skipped (--scale != 1 or --no-gate: no reference digests for this input)"
    if args.scale == 1.0 and not args.no_gate:
        gathered = [gate_rows]
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(gate_rows, gathered, dst=0)
        if rank == 0:
            golden = wl.load_digests()
            allrows = {}
            for g in gathered:
                allrows.update(g)
            bad = []
            for c in range(len(lengths)):
                keys = sorted(k for k in allrows if k[0] == c)
                assert keys, "contig %d was scanned by no rank" % c
                cat = lambda name: [np.concatenate([allrows[k][name][s] for k in keys]) for s in range(3)]  # noqa: E731
                lo, hi = wl.c3_window(c, lengths[c])
                want = golden["windows"][str(c)]
                assert (want["L"], want["lo"], want["hi"]) == (lengths[c], lo, hi), "golden made for another workload"
                win = cat("win")
                for s in range(3):
                    got = wl.digest_rows(win[s][:, 0], win[s][:, 1], win[s][:, 2])
                    if got != want["kept"][str(s + 1)]:
                        bad.append("contig %d window [%d,%d) stream %d: %s != reference %s" % (c, lo, hi, s + 1, got, want["kept"][str(s + 1)]))
                if c == golden["full"]["contig"]:
                    full = cat("full")
                    for s in range(3):
                        got = wl.digest_rows(full[s][:, 0], full[s][:, 1], full[s][:, 2])
                        if got != golden["full"]["kept"][str(s + 1)]:
                            bad.append("contig %d full length stream %d: %s != reference %s" % (c, s + 1, got, golden["full"]["kept"][str(s + 1)]))
            if bad:
                sys.stderr.write("bench.py: PARITY GATE FAILED\n" + "\n".join(bad) + "\n")
                sys.stderr.flush()
                os._exit(3)
            parity = "ok"
    gate_rows = None

    barrier()
    per_step, scan_ms, pack_ms, merge_ms, launches, restarts = [], [], [], [], 0, 0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                     # L2 flush between timed iterations
        torch.cuda.synchronize()
        tot = sc_ms = pk_ms = mg_ms = 0.0
        for sc, _, _ in ctxs:
            sc.scan_device()              # CUDA events on the library's stream bracket the kernels
            t = sc.timing()
            tot += t["total_ms"]; sc_ms += t["scan_ms"]; pk_ms += t["pack_ms"]; mg_ms += t["merge_ms"]
            launches += t["launches"]
            restarts += t["restarts"]
        per_step.append(tot); scan_ms.append(sc_ms); pack_ms.append(pk_ms); merge_ms.append(mg_ms)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    dev_ms = float(sum(per_step))
    dev_ms_max = allmax(dev_ms)
    dev_ms_min = -allmax(-dev_ms)
    genome = float(sum(lengths))
    value = genome * args.steps / (dev_ms_max * 1e-3) / 1e9
    counts_all = [int(allsum(c)) for c in counts]
    for sc, _, _ in ctxs:
        sc.close()
    ctxs = []

    # ---- host<->device copy ceiling of this box with all ranks copying at once (plain cudaMemcpyAsync, both directions) ------
    pc_n = 1 << 30
    pc_h = torch.empty(pc_n, dtype=torch.uint8, pin_memory=True)
    pc_h2 = torch.empty(pc_n, dtype=torch.uint8, pin_memory=True)
    pc_d = torch.empty(pc_n, dtype=torch.uint8, device="cuda")
    pc_d2 = torch.empty(pc_n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    pc_d.copy_(pc_h, non_blocking=True); pc_h2.copy_(pc_d2, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            pc_d.copy_(pc_h, non_blocking=True)
        with torch.cuda.stream(s2):
            pc_h2.copy_(pc_d2, non_blocking=True)
    torch.cuda.synchronize()
    pc_s = allmax(time.perf_counter() - t0)
    pcie_each = 3 * pc_n / pc_s / 1e9   # GB/s per direction per GPU, both directions busy, all ranks at once
    del pc_h, pc_h2, pc_d, pc_d2

    # ---- end to end through the public API with host buffers: every step copies its ASCII from pinned host memory
    # (rb_load_contigs), runs the kernels and copies its three streams back (rb_fetch_compact). Contigs go through
    # ribbit_b200.pipeline.ScanPipeline one by one: three contexts on the GPU, so the copies of one contig overlap the
    # kernels of the next. All results of all K steps are complete inside the timed region.
    # Whole contigs that lie next to each other in the host buffer go in batches of about --e2e-batch-mbp (the kernels run
    # at their large-batch rate); the smallest batch goes first (short pipeline fill), then the largest ones.
    units = [(g, None) for g in pipeline.group_contigs(sorted(whole), lengths, args.e2e_batch_mbp * 1_000_000)]
    units.sort(key=lambda u: -sum(lengths[c] for c in u[0]))
    if len(units) > 2:
        units.insert(0, units.pop())
    units += [([c], (a, b)) for c, a, b in ranges]
    trace = os.environ.get("RB_E2E_TRACE") == "1"  # diagnostic: rank 0 prints the pipeline timeline of the timed steps
    pipe = pipeline.ScanPipeline(M_LO, M_HI, device=local, depth=args.depth, compact=True, trace=trace)

    def submit_all():
        pipe.n = 0  # every step maps batch i to context i % depth: the contexts' buffers reach their final size in the warm-up
        return [pipe.submit_flat(host_np[offs[g[0]]:offs[g[-1]] + lengths[g[-1]] + 1], [lengths[c] for c in g], rng,
                                 offsets=[offs[c] - offs[g[0]] for c in g]) for g, rng in units]

    for _ in range(max(1, args.warmup)):
        for f in submit_all():
            f.result()
    barrier()
    if trace:
        pipe.trace.clear()
    t0 = time.perf_counter()
    e2e_counts = [0, 0, 0]
    futs = []
    for _ in range(args.steps):
        futs += submit_all()
    for k, f in enumerate(futs):
        res = f.result()
        if k >= len(futs) - len(units):
            for s in range(3):
                e2e_counts[s] += len(res[s][0])
    torch.cuda.synchronize()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = genome * args.steps / e2e_s / 1e9
    if trace:
        for k, nb, a, b_, c_, d_, e_ in sorted(pipe.trace, key=lambda x: x[2]):
            sys.stderr.write("rank %d " % rank + "ctx %d %9d bases: load %7.1f..%7.1f scan %7.1f..%7.1f fetch ..%7.1f ms\n" % (
                k, nb, (a - t0) * 1e3, (b_ - t0) * 1e3, (c_ - t0) * 1e3, (d_ - t0) * 1e3, (e_ - t0) * 1e3))
    if e2e_counts != counts:
        sys.stderr.write("bench.py: end-to-end streams differ from the device-resident ones: %s vs %s\n" % (e2e_counts, counts))
        os._exit(3)
    h2d = int(allsum(sum(lengths[c] for g, _ in units for c in g)))
    d2h = int(allsum(sum(e2e_counts) * 8 + sum(len(g) + 1 for g, _ in units) * 3 * 8))
    pipe.close()

    # ---- rank 0 only, outside the timed regions: the other BASELINE.json shapes (device-resident, few steps), the rows
    # next to the path (K0 rb_load_fasta, K7 rb_motif_rows) and the CPU baseline
    also, next_rows = None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    if extras:
        also = {}
        # C2 (configs[1]): one chr21-scale contig
        L2 = len(c2_seq)
        h2 = torch.empty(L2 + 64, dtype=torch.uint8, pin_memory=True)
        h2.numpy()[:L2] = np.frombuffer(c2_seq, dtype=np.uint8)
        d2 = h2.cuda()
        sc2 = scan.Scanner(M_LO, M_HI, device=local)
        sc2.load_device(d2.data_ptr(), [L2], keepalive=d2)
        ts = []
        for k in range(5):
            flush.zero_(); torch.cuda.synchronize()
            sc2.scan_device()
            ts.append(sc2.timing())
        tt = min(ts[2:], key=lambda t: t["total_ms"])
        also["c2"] = {"workload": "C2 chr21-scale contig, 46.7 Mbp, -m 2 -M 100", "gbp_per_s": L2 / tt["total_ms"] / 1e6,
                      "ms": tt["total_ms"], "scan_ms": tt["scan_ms"], "candidates": sc2.counts(),
                      "roofline_frac_int": ops_per_base(M_LO, M_HI) * L2 / (tt["scan_ms"] * 1e-3) / int_peak}
        # K0 / K7 on the same contig
        body = np.frombuffer(c2_seq, dtype=np.uint8)
        full = (L2 // 80) * 80
        lines = np.concatenate([body[:full].reshape(-1, 80), np.full((full // 80, 1), 10, np.uint8)], axis=1).reshape(-1)
        text_np = np.concatenate([np.frombuffer(b">chr21 synthetic\n", np.uint8), lines, body[full:], np.frombuffer(b"\n", np.uint8)])
        text = torch.empty(len(text_np), dtype=torch.uint8, pin_memory=True)
        text.numpy()[:] = text_np
        c2_counts = sc2.counts()
        t_fa = []
        for _ in range(4):
            t0 = time.perf_counter()
            names, lens = sc2.load_fasta(text.numpy())
            t_fa.append(time.perf_counter() - t0)
        assert names == ["chr21"] and lens.tolist() == [L2]
        sc2.scan_device()
        assert sc2.counts() == c2_counts, "streams differ after rb_load_fasta"
        a = sc2.fetch(copy=False)[2][0]
        a = a[(a["mlen"] > 10) & (a["flags"] == 0)][:400_000]
        seeds = np.stack([np.zeros(len(a), np.int32), a["start"], a["end"], a["mlen"].astype(np.int32)], axis=1).astype(np.int32)
        info = sc2.filter_seeds(seeds)
        seeds[:, 2] = np.minimum(seeds[:, 1] + info[:, 0], L2)
        t_mr = []
        for _ in range(3):
            t0 = time.perf_counter()
            rows = sc2.motif_rows(seeds)
            t_mr.append(time.perf_counter() - t0)
        next_rows = {"k0_load_fasta": {"text_bytes": int(len(text_np)), "ms": min(t_fa[1:]) * 1e3, "gbp_per_s": L2 / min(t_fa[1:]) / 1e9,
                                       "what": "rb_load_fasta from pinned host memory: H2D of the text + 3 kernels + header table D2H"},
                     "k7_motif_rows": {"seeds": int(len(seeds)), "ms": min(t_mr) * 1e3, "seeds_per_s": len(seeds) / min(t_mr),
                                       "scored": int((rows[:, 1] > 0).sum()),
                                       "what": "rb_motif_rows on the scan's kept anchored candidates with mlen > 10 (whole call: H2D of the seeds, kernel, D2H)"}}
        sc2.close()
        del d2, h2, text
        # C5 (configs[4]): short contigs, small-motif range
        n5 = len(c5_flat) // 1000
        h5 = torch.empty(len(c5_flat) + 64, dtype=torch.uint8, pin_memory=True)
        h5.numpy()[:len(c5_flat)] = np.frombuffer(c5_flat, dtype=np.uint8)
        d5 = h5.cuda()
        sc5 = scan.Scanner(1, 6, device=local)
        sc5.load_device(d5.data_ptr(), [1000] * n5, keepalive=d5)
        ts = []
        for k in range(5):
            flush.zero_(); torch.cuda.synchronize()
            sc5.scan_device()
            ts.append(sc5.timing())
        tt = min(ts[2:], key=lambda t: t["total_ms"])
        also["c5"] = {"workload": "C5 %d contigs x 1 kb, -m 1 -M 6" % n5, "gbp_per_s": len(c5_flat) / tt["total_ms"] / 1e6,
                      "ms": tt["total_ms"], "scan_ms": tt["scan_ms"], "candidates": sc5.counts(),
                      "roofline_frac_int": ops_per_base(1, 6) * len(c5_flat) / (tt["scan_ms"] * 1e-3) / int_peak,
                      "roofline_frac_hbm_ascii": 1.375 * len(c5_flat) / (tt["total_ms"] * 1e-3) / 1e9 / hbm_peak}
        sc5.close()
        del d5, h5

    if rank == 0:
        scan_ms_avg = float(np.mean(scan_ms))
        ops_per_launch = ops_per_base(M_LO, M_HI) * my_bases   # lane-operations (one 32-bit op in one lane = 32 bases x 1 op)
        achieved = ops_per_launch / (scan_ms_avg * 1e-3)
        hbm_achieved = PACKED_BYTES_PER_BASE * my_bases / (scan_ms_avg * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "scan_kernel_traffic.json"))).get("dram_bytes_per_launch_c3")
        except (OSError, ValueError):
            pass
        cpu = None
        if not args.no_extras:
            c0 = need[-1]
            cpu_v, cpu_kind, cpu_cores, cpu_what = cpu_scan_throughput(seqs[c0][:12_000_000], args.cpu_sample, min(ncpu, 8), "C3 contig %d" % c0)
            cpu = {"value": cpu_v, "unit": "Gbp/s", "cores": cpu_cores, "kind": cpu_kind, "sample": cpu_what}
        line = {
            "metric": "Gbp/s scanned (motifs 2-100)", "value": value, "unit": "Gbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "parity": parity,
            "config": {"workload": workload_name(args.scale), "genome_bases": int(genome), "contigs": len(lengths),
                       "min_mlen": M_LO, "max_mlen": M_HI,
                       "partition": {"how": "shard.plan_parts: whole contigs longest-first, then word ranges of the largest contigs (rb_set_word_range) until the loads differ by <= 2 %",
                                     "rank0_parts": [list(p) for p in mine], "split_contigs": sorted({c for p in plan for c, a, b in p if not (a == 0 and b == n_words[c])}),
                                     "rank_ms_min_max": [dev_ms_min / args.steps, dev_ms_max / args.steps]},
                       "l2": "flushed (512 MiB write) between timed steps; the inputs are far larger than L2",
                       "timing": "CUDA events on the library stream, summed over the rank's batches, max over ranks",
                       "candidates_per_step": counts_all,
                       "rank0_stage_ms": {"pack": float(np.mean(pack_ms)), "scan": scan_ms_avg, "merge": float(np.mean(merge_ms))},
                       "wall_ms_bracket": wall_ms, "warmup_restarts_per_step": restarts / args.steps,
                       "input_generation_s": gen_s, "also": also},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gbp/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "per step and batch: rb_load_contigs (pinned host ASCII -> HBM) + rb_scan_device (kernels) + rb_fetch_compact (D2H of the three streams as 8-byte records); batches of about %d Mbp pipelined over %d contexts per GPU, one batch's kernels at a time (ribbit_b200.pipeline)" % (args.e2e_batch_mbp, args.depth),
                    "pcie_gbs_per_direction": pcie_each,
                    "pcie_ceiling_gbps": world * min(pcie_each * 1e9 / (h2d / genome), pcie_each * 1e9 / (d2h / genome)) / 1e9 if h2d and d2h else None,
                    "pcie_note": "plain cudaMemcpyAsync of 1 GiB pinned buffers, H2D and D2H at the same time, all ranks at once; ceiling = that rate over the bytes per base each direction moves"},
            "gpu_launches": launches,
            "roofline": {"bound": "int", "kernel": "scan_kernel<32>", "achieved": achieved / 1e12, "peak": int_peak / 1e12,
                         "unit": "Tlaneop/s", "frac": achieved / int_peak if int_peak else None,
                         "note": "rank 0's share: algorithmic (12*NSHIFTS+39*NMOTIFS)/32 = %.1f word-ops per base x %d bases per launch / scan kernel time (CUDA events); peak = LOP3+SHF microbenchmark measured in this run; the HBM bound is ~100x looser" % (ops_per_base(M_LO, M_HI), my_bases),
                         "peak_variants_tlaneops": {k: v / 1e12 for k, v in int_peak_modes.items()} if int_peak_modes else None,
                         "traffic": traffic,
                         "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_achieved / hbm_peak, "peak_source": "measured" if peaks else "fallback"}},
            "cpu_baseline": cpu,
            "next_rows": next_rows,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale of the C3 contig lengths (1.0 = 3.1 Gbp; the parity gate needs 1.0)")
    ap.add_argument("--no-gate", action="store_true", help="skip the parity gate (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip C2 / C5 / K0 / K7 / CPU-baseline legs (profiling runs)")
    ap.add_argument("--depth", type=int, default=4, help="scan contexts per GPU in the end-to-end pipeline")
    ap.add_argument("--e2e-batch-mbp", type=int, default=400, help="bases per batch of the end-to-end pipeline (whole contigs)")
    ap.add_argument("--c5-contigs", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="bases per CPU-baseline process")
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
