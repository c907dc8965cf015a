"""Per-item timeline of the scan kernel (diagnostic). python tools/exp_items.py"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ribbit_b200 import scan, synth
seq = synth.contig_c2(46_700_000, seed=21)
L = len(seq)
h = torch.empty(L + 64, dtype=torch.uint8, pin_memory=True); h.numpy()[:L] = np.frombuffer(seq, dtype=np.uint8)
d = h.cuda()
sc = scan.Scanner(2, 100)
sc.load_device(d.data_ptr(), [L], keepalive=d)
sc.scan_device()
lib = sc.lib
lib.rb_debug_item_clocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
n = ctypes.c_int64()
lib.rb_debug_item_clocks(sc.ctx, None, 0, ctypes.byref(n))
sc.scan_device(); sc.scan_device()
print(sc.timing())
out = np.zeros((n.value, 4), np.int64)
assert lib.rb_debug_item_clocks(sc.ctx, out.ctypes.data, n.value, ctypes.byref(n)) == 0
ok = out[:, 1] > 0
t0 = out[ok, 0].min()
st = (out[:, 0] - t0) / 1e3; en = (out[:, 1] - t0) / 1e3
dur = en - st
nb = 4
print("items", n.value, "recorded", ok.sum(), "span us", en[ok].max())
for b in range(nb):
    m = ok & (np.arange(n.value) % nb == b)
    print("band", b, "dur us: mean %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f; start range %.0f..%.0f end max %.0f" % (
        dur[m].mean(), np.median(dur[m]), np.percentile(dur[m], 90), np.percentile(dur[m], 99), dur[m].max(), st[m].min(), st[m].max(), en[m].max()))
# concurrency over time
ev = np.concatenate([np.stack([st[ok], np.ones(ok.sum())], 1), np.stack([en[ok], -np.ones(ok.sum())], 1)])
ev = ev[np.argsort(ev[:, 0])]
conc = np.cumsum(ev[:, 1])
T = en[ok].max()
for f in np.linspace(0, 1, 21)[:-1]:
    i = np.searchsorted(ev[:, 0], f * T)
    print("t=%.0f us running items %d" % (f * T, conc[min(i, len(conc) - 1)]))
gen = out[:, 2]; rst = out[:, 3] >> 32; slw = out[:, 3] & 0xFFFFFFFF
print("general-path words per item: mean %.1f p50 %.0f p90 %.0f max %d; slow words mean %.2f; restarts mean %.2f" % (gen[ok].mean(), np.median(gen[ok]), np.percentile(gen[ok], 90), gen[ok].max(), slw[ok].mean(), rst[ok].mean()))
for lo, hi in ((0, 9), (9, 12), (12, 20), (20, 40), (40, 100), (100, 400), (400, 100000)):
    m = ok & (gen >= lo) & (gen < hi)
    if m.sum(): print("gen words [%d,%d): items %d mean dur %.0f us" % (lo, hi, m.sum(), dur[m].mean()))
A = np.stack([gen[ok], np.ones(ok.sum())], 1); coef = np.linalg.lstsq(A, dur[ok], rcond=None)[0]
print("fit: dur = %.2f us * general words + %.1f us" % (coef[0], coef[1]))
order = np.argsort(-np.where(ok, dur, 0))[:10]
nch = n.value // nb
cw_words = (L // 32 + nch - 1) // nch
for i in order:
    ch = i // nb
    print("item %d chunk %d band %d dur %.0f us start %.0f  ~words %d.. (pos %.2f Mbp) gen %d slow %d restarts %d" % (i, ch, i % nb, dur[i], st[i], ch * cw_words, ch * cw_words * 32 / 1e6, gen[i], slw[i], rst[i]))
