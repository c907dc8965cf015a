"""Attributes the per-SASS-instruction counters of an `ncu --page source --csv` dump to CUDA source lines, using the
line table of `nvdisasm -g -c` on the cubin of the same build (instructions are matched by order).

  python profiles/sass_by_line.py <ncu source csv> <nvdisasm -g -c listing> <mangled kernel name> [top]
"""
import csv
import re
import sys


def main():
    src_csv, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    inst = [r for r in rows[2:] if len(r) == len(hdr)]
    # disassembly lines of the kernel
    lines = open(dis).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kernel + ":"))
    loc = None
    locs = []
    for l in lines[start + 1:]:
        if l.startswith("//----") and ".text." in l:
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
            loc = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            locs.append(loc)
    n = min(len(locs), len(inst))
    agg = {}
    tot = 0
    for k in range(n):
        ie = int(inst[k][ix["Instructions Executed"]])
        te = int(inst[k][ix["Thread Instructions Executed"]])
        a = agg.setdefault(locs[k], [0, 0, 0])
        a[0] += ie; a[1] += te; a[2] += 1
        tot += ie
    print("instructions matched %d (sass rows %d, listing %d), warp instructions executed %d" % (n, len(inst), len(locs), tot))
    byfile = {}
    for (loc, v) in agg.items():
        f = loc[0] if loc else "?"
        byfile[f] = byfile.get(f, 0) + v[0]
    print("by file:", {k: "%.1f%%" % (100.0 * v / tot) for k, v in byfile.items()})
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%%  active %4.1f  sass %3d  %s" % (100.0 * v[0] / tot, v[1] / max(v[0], 1), v[2], loc))


if __name__ == "__main__":
    main()
