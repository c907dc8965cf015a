"""Host-side cost of the drop-in program up to the merged seed lists (before the per-seed stage) on the C2 contig."""
import sys, os, time, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribbit_b200 import synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 46_700_000
seq = synth.contig_c2(L, seed=21)
with tempfile.TemporaryDirectory() as td:
    fa = os.path.join(td, "x.fa"); synth.write_fasta(fa, [seq])
    for rep in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "ribbit_b200/bin/ribbit_gpu"), "-i", fa, "-o", os.path.join(td, "o.bed")],
                           env=dict(os.environ, RB_CP2_OUT="/dev/null", RB_CP_STOP_AFTER_CP2="1"), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        print("ribbit_gpu up to the merged seed lists: %.2f s (rc %d)" % (dt, r.returncode), "|", " ; ".join(l.split("\t")[0][:34] + " " + l.split("elapsed")[-1].strip(": ") for l in r.stderr.decode().split("\n") if "elapsed" in l), flush=True)
