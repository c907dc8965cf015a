"""Whole-program timing: the drop-in baseline/_ref/ribbit_gpu vs the unmodified reference oracle/_ref/ribbit_ref."""
import sys, os, time, subprocess, tempfile, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribbit_b200 import synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
seq = synth.contig_c2(L, seed=21)
with tempfile.TemporaryDirectory() as td:
    fa = os.path.join(td, "x.fa"); synth.write_fasta(fa, [seq])
    out = {}
    for name, exe in (("ribbit_gpu_nofilter", os.path.join(ROOT, "baseline/_ref/ribbit_gpu")), ("ribbit_gpu", os.path.join(ROOT, "baseline/_ref/ribbit_gpu")), ("ribbit_ref", os.path.join(ROOT, "oracle/_ref/ribbit_ref"))):
        bed = os.path.join(td, name + ".bed")
        t0 = time.perf_counter()
        env = dict(os.environ, RIBBIT_NO_SEED_FILTER="1") if name.endswith("nofilter") else dict(os.environ)
        env.pop("RIBBIT_NO_SEED_FILTER", None) if not name.endswith("nofilter") else None
        r = subprocess.run([exe, "-i", fa, "-o", bed], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        dt = time.perf_counter() - t0
        out[name] = (dt, hashlib.md5(open(bed, "rb").read()).hexdigest(), r.returncode)
        stages = [l for l in r.stderr.decode().split("\n") if "Time elapsed" in l]
        print(name, "%.2f s" % dt, "rc", r.returncode, "|", " ; ".join(s.split("\t")[0][:40] + " " + s.split("elapsed")[-1].strip(": ") for s in stages), flush=True)
    print("BED identical:", out["ribbit_gpu"][1] == out["ribbit_ref"][1] == out["ribbit_gpu_nofilter"][1], " speed-up %.2fx" % (out["ribbit_ref"][0] / out["ribbit_gpu"][0]))
