"""Join an ncu SASS-level source page with nvdisasm line info and aggregate stall samples / executed
instructions per CUDA source line.
usage: sass_hotspots.py <report.ncu-rep> <cubin> <kernel-mangled-substring> [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ilsb = hdr.index("stall_long_sb")
data = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# collect, per .text section, the ordered (opcode text, line) list
sections, cur, line = {}, None, None
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        cur = m.group(1); sections[cur] = []; line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur is not None:
        sections[cur].append((int(m.group(1), 16), m.group(2).strip(), line))
sec = [k for k in sections if kname in k]
assert sec, list(sections)[:5]
ins = sections[sec[0]]
print("kernel section", sec[0][:80], "nvdisasm instrs", len(ins), "ncu instrs", len(data))
base = int(data[0][ia], 16) if data[0][ia].startswith("0x") else int(data[0][ia])
by_line = defaultdict(lambda: [0, 0, 0])
tot_s = tot_e = 0
n = min(len(ins), len(data))
mism = 0
for k in range(n):
    addr, text, line = ins[k]
    r = data[k]
    if text.split()[0].strip("@!P0123456789") != r[isrc].split()[0].strip("@!P0123456789") and k < 50:
        mism += 1
    s, e, l = int(r[isamp] or 0), int(r[iex] or 0), int(r[ilsb] or 0)
    by_line[line][0] += s; by_line[line][1] += e; by_line[line][2] += l
    tot_s += s; tot_e += e
print("opcode mismatches in first 50:", mism, " total samples", tot_s, "executed", tot_e)
print("%-28s %8s %8s %8s" % ("file:line", "samples%", "exec%", "long_sb%"))
for line, (s, e, l) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %8.2f %8.2f %8.2f" % ("%s:%s" % line if line else "?", 100 * s / tot_s, 100 * e / tot_e, 100 * l / max(tot_s, 1)))
