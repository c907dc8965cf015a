import sys, numpy as np
import os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,ROOT+'/tests')
import oracle_util as ou, stream_model as sm
from ribbit_b200 import scan
g = np.load(ROOT+'/tests/golden/golden.npz')
names = sorted({k.rsplit('_',1)[0] for k in g.files})
import itertools
DBG=int(sys.argv[1]) if len(sys.argv)>1 else 0
for name in names:
    seq = g[name+'_seq'].tobytes(); mlo, mhi = [int(x) for x in g[name+'_args']]
    exp = sm.expected_streams(seq, ou.scan_events(seq, mlo, mhi))
    for cw in (0, 5, 33):
        sc = scan.Scanner(mlo, mhi, chunk_words=cw, debug=DBG); sc.load([seq]); got = scan.contig_streams(sc.scan(), 0); sc.close()
        for s in (1,2,3):
            if got[s].shape != exp[s].shape or not (got[s]==exp[s]).all():
                print('MISMATCH dbg', DBG, name, 'cw', cw, 'stream', s, got[s].shape, exp[s].shape)
                n = min(len(got[s]), len(exp[s]))
                d = np.flatnonzero((got[s][:n] != exp[s][:n]).any(axis=1))
                i = d[0] if len(d) else n
                print(' first diff at', i, ' got', got[s][i:i+1].tolist(), ' exp', exp[s][i:i+1].tolist())
                break
print('done')
