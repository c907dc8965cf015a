import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
from ribbit_b200 import scan, synth
seq = synth.contig_c2(10_000_000, seed=21)
t1 = time.perf_counter()
sc = scan.Scanner(2, 100); t2 = time.perf_counter()
sc.load([seq]); t3 = time.perf_counter()
r = sc.scan(copy=False); t4 = time.perf_counter()
h = sc.planes(0); t5 = time.perf_counter()
a = sc.anchor_planes(0, 1, 102); t6 = time.perf_counter()
sc.load([seq]); r = sc.scan(copy=False); t7 = time.perf_counter()
print("import+gen %.2f s | rb_create %.3f s | first load %.3f s | first scan+fetch %.3f s | planes %.3f s | anchor planes (102 x 10 Mbp) %.3f s | second load+scan %.3f s" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))
