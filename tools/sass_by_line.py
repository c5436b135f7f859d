#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo): a quick way to see what a
source change did to the code the compiler emitted, without a GPU.

    python tools/sass_by_line.py gr_adsb_b200/csrc/libadsb_b200.so 'scan_slice_kernelILi2ELi1E' [top_n]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
        cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
        out = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True, check=True).stdout
    cur_fn, cur_line = None, None
    per_line = collections.Counter()
    total = 0
    for ln in out.splitlines():
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if cur_fn and pat in cur_fn and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", ln):
            per_line[cur_line] += 1
            total += 1
    print("total", total)
    for (k, v) in per_line.most_common(top):
        print("%5d  %s:%s" % (v, k[0] if k else "?", k[1] if k else "?"))


if __name__ == "__main__":
    main()
