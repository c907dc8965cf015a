"""Per-function summary of an `ncu --page source --csv` export of a kernel with out-of-line device functions.
   python tools/ncu_fn.py <src.csv> <lib.so> <kernel substring> <steps>
Functions are matched by their order and size in the kernel's .text section (nvdisasm)."""
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
print("csv instructions", len(inst), "sass instructions", sum(f[1] for f in funcs))
k = 0
tot_s = sum(int(r[ix['# Samples']]) for r in inst)
for name, n in funcs:
    seg = inst[k:k + n]; k += n
    ie = sum(int(r[ix['Instructions Executed']]) for r in seg)
    sm = sum(int(r[ix['# Samples']]) for r in seg)
    short = re.sub(r"^.*\$", "", name)[:60]
    if ie: print("%-62s n=%5d  exec %10.3e = %7.1f/step  samples %5.1f%%" % (short, n, ie, ie / steps, 100.0 * sm / max(tot_s, 1)))
