"""Dynamic op-class histogram per function of a kernel from an `ncu --page source --csv` export.
   python profiles/opmix.py <src.csv> <lib.so> <kernel substring> <word steps (words x bands)>
Functions are matched by their order and size in the kernel's .text section (nvdisasm); counts are warp instructions
executed per word step of the whole launch."""
import csv, re, subprocess, sys, tempfile, os, collections
src, lib, kern, steps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
td = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null && for f in *.cubin; do nvdisasm -c $f; done > dis.txt" % (td, os.path.abspath(lib)), shell=True, check=True)
lines = open(os.path.join(td, "dis.txt")).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.endswith(":"))
funcs = [[lines[start][6:-1], 0]]
for l in lines[start + 1:]:
    if l.startswith("//----") and ".text." in l: break
    m = re.match(r"^([\$_A-Za-z][\w\$]*):$", l)
    if m and not l.startswith(".L"):
        funcs.append([m.group(1), 0]); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l): funcs[-1][1] += 1
rows = list(csv.reader(open(src)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
inst = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[ix['Instructions Executed']].isdigit()]
CLASSES = ["LOP3", "SHF", "IMAD.MOV", "MOV", "IMAD", "IADD3", "VIADD", "LEA", "ISETP", "SEL", "PRMT", "PLOP3", "BRA", "BSSY", "BSYNC", "VOTE", "SHFL", "REDUX", "FLO", "BREV", "POPC", "LDG", "LD", "ST", "STG", "LDL", "STL", "UMOV", "CALL", "RET"]
print("csv instructions %d, sass instructions %d, word steps %.4g" % (len(inst), sum(f[1] for f in funcs), steps))
k = 0
for name, n in funcs:
    seg = inst[k:k + n]; k += n
    ops = collections.Counter(); tot = 0.0; thr = 0.0
    for r in seg:
        e = int(r[ix['Instructions Executed']]) / steps
        s = r[ix['Source']].strip().split()
        op = s[1] if s[0].startswith("@") else s[0]
        base = "IMAD.MOV" if op.startswith("IMAD.MOV") else op.split(".")[0]
        ops[base] += e; tot += e; thr += int(r[ix['Thread Instructions Executed']]) / steps
    if tot < 0.05: continue
    short = re.sub(r"^.*\$", "", name)[:70]
    print("\n%s: %.1f warp instructions per word step of the launch, %.1f active lanes" % (short, tot, thr / max(tot, 1e-9)))
    print("  " + ", ".join("%s %.1f" % (c, ops[c]) for c in CLASSES if ops[c] >= 0.05))
    rest = sum(v for c, v in ops.items() if c not in CLASSES)
    if rest >= 0.05: print("  other %.1f" % rest)
