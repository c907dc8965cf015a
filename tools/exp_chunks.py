"""scan time of the C2 contig vs chunk size.  python tools/exp_chunks.py [mbp]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ribbit_b200 import scan, synth
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 46.7
seq = synth.contig_c2(int(mbp * 1e6), seed=21)
L = len(seq)
h = torch.empty(L + 64, dtype=torch.uint8, pin_memory=True)
h.numpy()[:L] = np.frombuffer(seq, dtype=np.uint8)
d = h.cuda()
for cw in (0, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 2048):
    sc = scan.Scanner(2, 100, chunk_words=cw)
    sc.load_device(d.data_ptr(), [L], keepalive=d)
    ts = []
    for _ in range(4):
        sc.scan_device(); ts.append(sc.timing()["scan_ms"])
    print("chunk_words", cw, "scan_ms", min(ts), "restarts", sc.timing()["restarts"], flush=True)
    sc.close()
