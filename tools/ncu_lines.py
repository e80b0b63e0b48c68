#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_REGEX [top] -- per-SOURCE-LINE profile of one captured kernel.

`ncu --page source --csv` gives stall samples and executed-instruction counts per SASS instruction but (in CSV form) no source
correlation; `nvdisasm -g` gives the source line (and inlining chain) of every SASS instruction of the cubin inside LIB.so.  The
two listings are joined by instruction order.  Output: share of warp-stall samples and of executed warp instructions per source
line (innermost frame), sorted by samples -- the table profiles/*_lines.txt hold."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel_re):
    """[(sass text, 'file:line <- file:line ...')] of the first kernel in LIB matching kernel_re."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        inside, cur = False, ""
        for line in txt.splitlines():
            if line.startswith("//--------------------- .text."):
                if inside and out:
                    return out
                inside = re.search(kernel_re, line) is not None
                continue
            if not inside:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
            if m:
                chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                cur = " <- ".join(["%s:%s" % (os.path.basename(m.group(1)), m.group(2))] +
                                  ["%s:%s" % (os.path.basename(a), b) for a, b in chain])
                continue
            m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(.*?);', line)
            if m:
                out.append((m.group(1).strip(), cur))
        if out:
            return out
    return out


def main():
    rep, lib, kre = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if "# Samples" in r)
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    sl = sass_lines(lib, kre)
    if len(sl) != len(data):
        print("# WARNING: %d SASS instructions in the report vs %d in %s -- was the library rebuilt since the capture?" %
              (len(data), len(sl), lib))
    n = min(len(sl), len(data))
    agg = {}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for k in range(n):
        r = data[k]
        key = sl[k][1]
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += int(r[col["# Samples"]] or 0)
        a[1] += int(r[col["Instructions Executed"]] or 0)
        for c in stall_cols:
            a[2][c[6:]] = a[2].get(c[6:], 0) + int(r[col[c]] or 0)
    S = sum(a[0] for a in agg.values()) or 1
    N = sum(a[1] for a in agg.values()) or 1
    print("# %s  kernel /%s/  %d SASS instructions, %d samples, %d executed warp instructions" % (os.path.basename(rep), kre, n, S, N))
    print("# samples%%  inst%%   top stalls                       source line (innermost <- inlined at ...)")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ", ".join("%s %d%%" % (k, 100 * v // max(a[0], 1)) for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:2])
        print("%8.2f %7.2f   %-32s %s" % (100.0 * a[0] / S, 100.0 * a[1] / N, st, key))


if __name__ == "__main__":
    main()
