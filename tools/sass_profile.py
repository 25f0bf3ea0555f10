"""Aggregate an ncu `--page source --csv` dump by source line (using nvdisasm --print-line-info on the .so)
to see where the executed instructions of a kernel go.  Usage:
    python tools/sass_profile.py <report.ncu-rep> <mangled kernel substring> [lib.so]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(lib, kern):
    cubins = subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd="/tmp", capture_output=True, text=True)
    out = []
    for f in re.findall(r"Extracting ELF file\s+\d+:\s+(\S+)", cubins.stdout):
        txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join("/tmp", f)], capture_output=True, text=True).stdout
        lines = txt.split("\n")
        starts = [i for i, l in enumerate(lines) if l.strip().startswith(".text.") and kern in l]
        if not starts:
            continue
        s = starts[0]
        e = next((i for i in range(s + 1, len(lines)) if lines[i].strip().startswith(".text.")), len(lines))
        cur = None
        for l in lines[s:e]:
            m = re.search(r'//## File "(.*?)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                out.append(cur)
        break
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "cineform-sdk_b200", "libcfhd_b200.so")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    data = rows[2:]
    ia, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    lm = line_map(lib, kern)
    print(f"{len(data)} SASS instructions in report, {len(lm)} in line map")
    by_line, by_op, stall_line = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for i, r in enumerate(data):
        n = int(r[ia])
        tot += n
        src = r[isrc].strip()
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        by_op[op] += n
        key = lm[i] if i < len(lm) and lm[i] else ("?", 0)
        by_line[key] += n
        stall_line[key] += int(r[ist])
    print("total warp instructions", tot)
    print("-- by opcode")
    for op, n in by_op.most_common(18):
        print(f"  {op:10s} {n / tot * 100:5.1f}%")
    print("-- by source line (file:line  share  stall-samples)")
    src_cache = {}
    for (f, ln), n in by_line.most_common(45):
        path = os.path.join(ROOT, "cineform-sdk_b200", "csrc", f)
        if path not in src_cache and os.path.exists(path):
            src_cache[path] = open(path).read().split("\n")
        text = src_cache.get(path, [""] * (ln + 1))[ln - 1].strip()[:90] if ln else ""
        print(f"  {f}:{ln:<5d} {n / tot * 100:5.1f}%  {stall_line[(f, ln)]:6d}  {text}")


if __name__ == "__main__":
    main()
