"""SASS instruction count per source line range of one kernel (needs -lineinfo): where the code bytes of a kernel come from.

    python tools/sass_lines.py textreid_b200/build/loss_fused.o fused_loss_kernel [bucket=10] [min=100]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, kern = sys.argv[1], sys.argv[2]
    bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    floor = int(sys.argv[4]) if len(sys.argv) > 4 else 100
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    cnt = collections.Counter()
    func = line = None
    for l in sass.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            func = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)) // bucket * bucket)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l) and func and kern in func:
            cnt[line] += 1
    print("total", sum(cnt.values()), "instructions =", sum(cnt.values()) * 16 // 1024, "KB")
    for k, c in sorted(cnt.items()):
        if c >= floor:
            print("%-24s %5d  %d" % (k[0], k[1], c))


if __name__ == "__main__":
    main()
