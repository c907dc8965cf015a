"""Per-item timeline of the scan kernel on the C5 shape (diagnostic). python tools/exp_items_c5.py [n_contigs]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ribbit_b200 import scan, synth
n_c = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
flat = b"".join(synth.contigs_c5(n=n_c, length=1000, seed=5))
h = torch.empty(len(flat) + 64, dtype=torch.uint8, pin_memory=True)
h.numpy()[:len(flat)] = np.frombuffer(flat, dtype=np.uint8)
d = h.cuda()
sc = scan.Scanner(1, 6)
sc.load_device(d.data_ptr(), [1000] * n_c, keepalive=d)
sc.scan_device()
lib = sc.lib
lib.rb_debug_item_clocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
n = ctypes.c_int64()
lib.rb_debug_item_clocks(sc.ctx, None, 0, ctypes.byref(n))
sc.scan_device(); sc.scan_device()
print(sc.timing(), sc.counts())
out = np.zeros((n.value, 4), np.int64)
assert lib.rb_debug_item_clocks(sc.ctx, out.ctypes.data, n.value, ctypes.byref(n)) == 0
ok = out[:, 1] > 0
t0 = out[ok, 0].min()
st = (out[:, 0] - t0) / 1e3; en = (out[:, 1] - t0) / 1e3
dur = en - st
print("items", n.value, "recorded", ok.sum(), "span us", en[ok].max())
print("dur us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (dur[ok].mean(), np.median(dur[ok]), np.percentile(dur[ok], 90), np.percentile(dur[ok], 99), dur[ok].max()))
gen = out[:, 2]; rst = out[:, 3] >> 32; slw = out[:, 3] & 0xFFFFFFFF
print("general-path words per item: mean %.1f p50 %.0f p90 %.0f max %d; slow words mean %.2f; restarts mean %.2f" % (gen[ok].mean(), np.median(gen[ok]), np.percentile(gen[ok], 90), gen[ok].max(), slw[ok].mean(), rst[ok].mean()))
A = np.stack([gen[ok], slw[ok], np.ones(ok.sum())], 1); coef = np.linalg.lstsq(A, dur[ok], rcond=None)[0]
print("fit: dur = %.2f us * general words + %.2f us * slow words + %.1f us" % tuple(coef))
ev = np.concatenate([np.stack([st[ok], np.ones(ok.sum())], 1), np.stack([en[ok], -np.ones(ok.sum())], 1)])
ev = ev[np.argsort(ev[:, 0])]
conc = np.cumsum(ev[:, 1])
T = en[ok].max()
for f in np.linspace(0, 1, 11)[:-1]:
    i = np.searchsorted(ev[:, 0], f * T)
    print("t=%.0f us running items %d" % (f * T, conc[min(i, len(conc) - 1)]))
sc.close()
