"""Whole-program timing: the drop-in ribbit_b200/bin/ribbit_gpu vs the unmodified reference oracle/_ref/ribbit_ref.
usage: python tools/exp_cli.py [bases] [c1|c2]; ribbit_gpu_hostmotif (make -C ribbit_b200/host HOST_MOTIF=1 BIN=...) = the drop-in
without K7, if it was built."""
import sys, os, time, subprocess, tempfile, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ribbit_b200 import synth
L = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
shape = sys.argv[2] if len(sys.argv) > 2 else 'c2'
seq = synth.contig_c2(L, seed=21) if shape == 'c2' else synth.contig_c1(L, seed=21)
with tempfile.TemporaryDirectory() as td:
    fa = os.path.join(td, "x.fa"); synth.write_fasta(fa, [seq])
    out = {}
    arms = [("ribbit_gpu_nofilter", os.path.join(ROOT, "ribbit_b200/bin/ribbit_gpu")), ("ribbit_gpu", os.path.join(ROOT, "ribbit_b200/bin/ribbit_gpu"))]
    if os.environ.get("EXP_CLI_SKIP_NOFILTER"):
        arms = arms[1:]
    if os.path.exists(os.path.join(ROOT, "ribbit_b200/bin/ribbit_gpu_hostmotif")):
        arms.append(("ribbit_gpu_hostmotif", os.path.join(ROOT, "ribbit_b200/bin/ribbit_gpu_hostmotif")))
    arms.append(("ribbit_ref", os.path.join(ROOT, "oracle/_ref/ribbit_ref")))
    for name, exe in arms:
        bed = os.path.join(td, name + ".bed")
        t0 = time.perf_counter()
        env = dict(os.environ, RIBBIT_NO_SEED_FILTER="1") if name.endswith("nofilter") else dict(os.environ, RIBBIT_VERBOSE="1")
        env.pop("RIBBIT_NO_SEED_FILTER", None) if not name.endswith("nofilter") else None
        r = subprocess.run([exe, "-i", fa, "-o", bed], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
        dt = time.perf_counter() - t0
        out[name] = (dt, hashlib.md5(open(bed, "rb").read()).hexdigest(), r.returncode)
        stages = [l for l in r.stderr.decode().split("\n") if "Time elapsed" in l]
        print("\n".join(l for l in r.stderr.decode().split("\n") if l.startswith("K7")))
        print(name, "%.2f s" % dt, "rc", r.returncode, "|", " ; ".join(s.split("\t")[0][:40] + " " + s.split("elapsed")[-1].strip(": ") for s in stages), flush=True)
    print("BED identical:", len({v[1] for v in out.values()}) == 1, " speed-up %.2fx" % (out["ribbit_ref"][0] / out["ribbit_gpu"][0]))
