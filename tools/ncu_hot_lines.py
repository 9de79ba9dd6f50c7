"""Warp-stall samples of one ncu capture aggregated PER SOURCE LINE.

    python tools/ncu_hot_lines.py <report.ncu-rep> <library.so> [top_n]

`ncu -i report --page source --csv` lists every SASS instruction of the captured kernel with its stall samples;
`nvdisasm -g` on the cubin extracted from the library (built with -lineinfo) maps instruction offsets to file:line.
Joined on the offset inside the kernel, summed per line, printed with the dominant stall reasons.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def norm(name):
    name = re.sub(r"\((?:bool|int|unsigned int|long)\)", "", name).replace(" ", "")
    return name.replace("false", "0").replace("true", "1")


def main():
    rep, lib = sys.argv[1], os.path.abspath(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    lines = out.splitlines()
    kernel = next(csv.reader([lines[0]]))[1]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
    base = int(rows[0]["Address"], 16)
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    # mangled name: template arguments of the demangled name decide which cubin function to disassemble
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
        sass = ""
        for cb in cubins:
            sass += subprocess.run(["nvdisasm", "-g", "-c", cb], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    # split per function, pick the one whose demangled name matches
    funcs, cur, name = {}, None, None
    for ln in sass.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"//-+ \.text\.(\S+) -+", ln)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            continue
        if cur is not None:
            cur.append(ln)
    want = None
    for mangled in funcs:
        dem = subprocess.run(["c++filt", mangled], stdout=subprocess.PIPE, text=True).stdout.strip()
        if norm(dem) == norm(kernel):
            want = mangled
    if want is None:
        sys.exit("kernel %r not found among %d functions" % (kernel, len(funcs)))
    off2line, where = {}, "?"
    for ln in funcs[want]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            where = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
            inl = re.search(r'inlined at "([^"]+)", line (\d+)', ln)
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            off2line[int(m.group(1), 16)] = where
    agg = collections.defaultdict(lambda: collections.Counter())
    total = 0
    for r in rows:
        n = int(r["# Samples"] or 0)
        if not n:
            continue
        total += n
        w = off2line.get(int(r["Address"], 16) - base, "?")
        agg[w]["_n"] += n
        agg[w]["_inst"] += int(r["Instructions Executed"] or 0)
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                agg[w][c] += v
    print("%s\n%d stall samples; top %d source lines" % (kernel, total, top))
    for w, c in sorted(agg.items(), key=lambda kv: -kv[1]["_n"])[:top]:
        reasons = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / c["_n"]) for k, v in c.most_common(5) if not k.startswith("_"))
        print("%6.2f%%  %-28s inst %-10d %s" % (100.0 * c["_n"] / total, w, c["_inst"], reasons))


main()
