#!/usr/bin/env python
"""Attribute ncu per-SASS executed-instruction counts to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <kernel-regex> <mangled-substring> [top]
Needs the .so built with -lineinfo; uses cuobjdump/nvdisasm here (no GPU).
"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kregex, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "advancedps.jl_b200", "libaps_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {so} > /dev/null && nvdisasm -g -c *.cubin > dis.txt 2>/dev/null", shell=True, check=True)
addr2line = {}
cur, infn = None, False
for ln in open(os.path.join(tmp, "dis.txt")):
    if ln.startswith("//---") and ".text." in ln:
        infn = mangled in ln
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(data[0][0], 16)
# the report may hold several launches of the kernel back to back: fold by offset
nfun = len(addr2line)
byline, bysamp, tot, launches = defaultdict(float), defaultdict(float), 0, max(1, len(data) // max(nfun, 1))
for k, r in enumerate(data):
    off = (k % nfun) * 16
    n = int(r[ix["Instructions Executed"]])
    s = int(r[ix["# Samples"]])
    key = addr2line.get(off, (None, ""))[0]
    byline[key] += n
    bysamp[key] += s
    tot += n
grid = None
print(f"launches folded: {launches}, SASS instrs: {nfun}, total warp-instr per launch: {tot/launches:.0f}")
src_cache = {}
def src(key):
    if not key: return ""
    f, l = key
    for d in ("advancedps.jl_b200/csrc", "include"):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().split("\n")
            return src_cache[p][l - 1].strip()[:80]
    return ""
stot = sum(bysamp.values()) or 1
for key, n in sorted(byline.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{str(key):32s} {100*n/tot:5.1f}% inst  {100*bysamp[key]/stot:5.1f}% stall-samples  {src(key)}")
