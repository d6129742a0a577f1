"""Executed warp instructions of one kernel by CUDA source line, from an ncu report captured with --import-source on:
joins ncu's SASS page (per-instruction counts) with nvdisasm's line info of the same cubin by instruction order.
usage: python profiles/sass_lines.py <report.ncu-rep> <kernel-regex> <cubin-name-part> [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre, cub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
sti = hdr.index("Warp Stall Sampling (All Samples)")
ncu_rows = [(r[si].strip(), int(r[ci]), int(r[ti]), int(r[sti] or 0)) for r in rows[hi + 1:] if len(r) > ci and r[ci].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "hive-alphazero_b200", "lib", "libhive_b200.so")], cwd=tmp, capture_output=True)
cubin = next(os.path.join(tmp, f) for f in os.listdir(tmp) if cub in f)
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and re.search(kre, l))
seq, cur = [], ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(1).strip(), cur))
print("ncu instructions", len(ncu_rows), "nvdisasm instructions", len(seq))
n = min(len(ncu_rows), len(seq))
by_line = collections.Counter(); thr = collections.Counter(); stall = collections.Counter()
for i in range(n):
    by_line[seq[i][1]] += ncu_rows[i][1]; thr[seq[i][1]] += ncu_rows[i][2]; stall[seq[i][1]] += ncu_rows[i][3]
total = sum(by_line.values()); stot = max(sum(stall.values()), 1)
print("total warp instructions", total)
src = {}
for (f, ln), c in by_line.most_common(top):
    if f not in src:
        p = os.path.join(ROOT, "hive-alphazero_b200", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print("%9d %5.1f%%  lanes %4.1f  stall %4.1f%%  %s:%d  %s" % (c, 100.0 * c / total, thr[(f, ln)] / max(c, 1), 100.0 * stall[(f, ln)] / stot, f, ln, text))

# executed instructions between consecutive block barriers in SASS layout order (= the phases of hive_step_kernel)
seg, acc, accs, k = [], 0, 0, 0
for i in range(n):
    acc += ncu_rows[i][1]; accs += ncu_rows[i][3]
    if seq[i][0].startswith("BAR.SYNC") or i == n - 1:
        seg.append((k, acc, accs)); acc = 0; accs = 0; k += 1
print("segments split at BAR.SYNC (layout order):")
for k, c, st in seg:
    print("   segment %d: %9d warp instructions (%4.1f%%), stall samples %4.1f%%" % (k, c, 100.0 * c / total, 100.0 * st / stot))
