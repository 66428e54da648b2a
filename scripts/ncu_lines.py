"""Per-source-line instruction counts and stall samples of one launch of an .ncu-rep.
usage: python scripts/ncu_lines.py rep.ncu-rep LAUNCH_INDEX cubin KERNEL_SUBSTR [top]
Joins ncu's per-SASS-instruction page (--page source --print-source sass) with nvdisasm's line info of the cubin."""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, idx, cubin, ksub = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(int(r[ia], 16), r[1].strip(), int(r[ii] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > ii and r[ia].startswith("0x")]
base = ins[0][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = ksub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for addr, txt, n, s in ins:
    k = line_of.get(addr - base)
    agg[k][0] += n
    agg[k][1] += s
    tot_i += n
    tot_s += s
print(f"kernel {rows[0][1][:80]}: {tot_i} warp instructions, {tot_s} samples")
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*s/max(tot_s,1):6.2f}% samples {100*n/max(tot_i,1):6.2f}% inst  {k}")
