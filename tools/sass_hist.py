"""Op-class histogram of the loops of a function in the built library.
   python tools/sass_hist.py <lib.so> <mangled function substring> [min loop size]"""
import re, subprocess, sys, tempfile, os, collections
lib, kern = sys.argv[1], sys.argv[2]
minn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
td = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null && for f in *.cubin; do nvdisasm -c $f; done > dis.txt" % (td, os.path.abspath(lib)), shell=True, check=True)
lines = open(os.path.join(td, "dis.txt")).read().split("\n")
starts = [i for i, l in enumerate(lines) if (l.startswith(".text.") or re.match(r"^\s*\.type\s", l) is None and re.match(r"^[_A-Za-z$][\w$]*:$", l)) and kern in l and l.endswith(":")]
start = starts[0]
print("function:", lines[start])
body = []
for l in lines[start + 1:]:
    if l.startswith("//----") and ".text." in l: break
    if re.match(r"^[_A-Za-z$][\w$]*:$", l) and not l.startswith(".L"): break
    body.append(l)
labels = {}
ins = []
for l in body:
    m = re.match(r"^(\.L_x_\d+):", l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append(m.group(2))
def opname(s):
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    base = op.split(".")[0]
    if base == "IMAD" and (".MOV" in op): return "IMAD.MOV"
    if base == "IMAD" and (".SHL" in op): return "IMAD.SHL"
    return base
print("instructions:", len(ins))
loops = []
for i, x in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+.*`\((\.L_x_\d+)\)", x)
    if m and m.group(1) in labels and labels[m.group(1)] <= i:
        loops.append((labels[m.group(1)], i))
ALU = {"LOP3", "SHF", "IADD3", "ISETP", "SEL", "LEA", "PRMT", "VIADD", "IADD", "PLOP3", "VIMNMX", "VIADDMNMX", "IABS", "MOV", "P2R", "R2P", "FLO", "BREV", "POPC"}
for a, b in sorted(set(loops), key=lambda t: t[1] - t[0], reverse=True):
    seg = ins[a:b + 1]
    if len(seg) < minn: continue
    ops = collections.Counter(opname(s) for s in seg)
    print("loop %5d..%5d n=%4d : %s" % (a, b, len(seg), ", ".join("%s:%d" % kv for kv in ops.most_common(40))))
