"""Stage times of the drop-in program (RIBBIT_VERBOSE=1) on a C2-shape contig.   python tools/exp_stages.py [bases]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribbit_b200 import synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 46_700_000
with tempfile.TemporaryDirectory() as td:
    fa = os.path.join(td, "x.fa"); synth.write_fasta(fa, [synth.contig_c2(L, seed=21)])
    for rep in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "ribbit_b200/bin/ribbit_gpu"), "-i", fa, "-o", os.path.join(td, "o.bed")],
                           env=dict(os.environ, RIBBIT_VERBOSE="1"), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        print("run %d: %.2f s total, rc %d" % (rep, dt, r.returncode))
        print("\n".join(l for l in r.stderr.decode().split("\n") if "[stage]" in l or l.startswith("K7")))
