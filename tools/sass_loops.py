"""Lists the loops (backward branches) of a kernel in an `nvdisasm -c` listing with instruction / MOV / branch counts.
   python tools/sass_loops.py <lib.so> <mangled kernel substring>"""
import re, subprocess, sys, tempfile, os, collections
lib, kern = sys.argv[1], sys.argv[2]
td = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null && for f in *.cubin; do nvdisasm -c $f; done > dis.txt" % (td, os.path.abspath(lib)), shell=True, check=True)
lines = open(os.path.join(td, "dis.txt")).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.endswith(":"))
body = []
for l in lines[start + 1:]:
    if l.startswith("//----") and ".text." in l: break
    body.append(l)
labels = {}
ins = []
for l in body:
    m = re.match(r"^(\.L_x_\d+):", l)
    if m: labels[m.group(1)] = len(ins); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append(m.group(2))
print("kernel instructions:", len(ins), "MOV-like:", sum(1 for x in ins if "MOV" in x.split()[0] or (x.startswith("@") and "MOV" in x)))
loops = []
for i, x in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+.*`\((\.L_x_\d+)\)", x)
    if m and m.group(1) in labels and labels[m.group(1)] <= i:
        loops.append((labels[m.group(1)], i))
for a, b in sorted(set(loops), key=lambda t: t[1] - t[0], reverse=True):
    seg = ins[a:b + 1]
    ops = collections.Counter((s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0] for s in seg)
    print("loop %5d..%5d  n=%4d  MOV=%3d BRA=%3d SHFL=%d REDUX=%d LDG=%d STG=%d ATOMS=%d  top: %s" % (
        a, b, len(seg), ops["IMAD"] and sum(1 for s in seg if "IMAD.MOV" in s) + ops["MOV"], ops["BRA"], ops["SHFL"], ops["REDUX"], ops["LDG"], ops["STG"], ops["ATOMS"],
        ", ".join("%s:%d" % kv for kv in ops.most_common(6))))
