#!/usr/bin/env python
"""tools/sass_fingerprint.py LIB.so [OUT.json] -- per-kernel fingerprint (instruction count + md5 of the instruction text) of every
kernel in a built library, with the file-hash part of anonymous-namespace names removed.  Two builds whose fingerprints agree for a
kernel run the same SASS for it: the way to show that a host-side refactoring (e.g. making a device function host-testable) left
the device code untouched without spending GPU time.  With two JSON files: prints the kernels that differ."""
import hashlib
import json
import re
import subprocess
import sys


def fingerprint(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, h, n = {}, None, None, 0
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out[name] = [n, h.hexdigest()]
            name = re.sub(r"_GLOBAL__N__[0-9a-f]{8}_(\d+)_(\w+?)_cu_[0-9a-f]{8}", r"_GLOBAL__N__\2_cu", m.group(1))
            h, n = hashlib.md5(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?;)", line)
        if m and name:
            h.update(m.group(1).encode())
            n += 1
    if name:
        out[name] = [n, h.hexdigest()]
    return out


if __name__ == "__main__":
    if sys.argv[1].endswith(".json"):
        a, b = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
        diff = sorted(k for k in set(a) | set(b) if a.get(k) != b.get(k))
        for k in diff:
            print(k[:140], a.get(k), "->", b.get(k))
        print(f"{len(diff)} of {len(set(a) | set(b))} kernels differ")
    else:
        fp = fingerprint(sys.argv[1])
        if len(sys.argv) > 2:
            json.dump(fp, open(sys.argv[2], "w"), indent=0, sort_keys=True)
        print(len(fp), "kernels")
