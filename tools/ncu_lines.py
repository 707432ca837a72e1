"""Per-source-line instruction / stall summary of one kernel in an .ncu-rep.

ncu's CSV source page is SASS-level; nvdisasm --print-line-info on the cubin of the shipped .so
gives the line of every SASS offset.  The two are joined on the offset inside the kernel.
usage: python tools/ncu_lines.py <rep> <mangled-name-substring> [top N] [ncu kernel regex]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ncu_kern = sys.argv[4] if len(sys.argv) > 4 else kern
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pdsat_b200", "libpdsat_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
line_of = {}
cur, infunc = None, False
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        infunc = kern in l
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + ncu_kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ia, ii, it, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > it and r[ia].startswith("0x")]
base = int(body[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in body:
    off = int(r[ia], 16) - base
    ln = line_of.get(off, (("?", 0), ""))[0]
    v = (int(r[ii]), int(r[it]), int(r[iss]))
    for j in range(3):
        agg[ln][j] += v[j]
        tot[j] += v[j]
src = {}
print("total warp-instr %d  thread-instr %d  (%.1f lanes)  samples %d" % (tot[0], tot[1], tot[1] / max(1, tot[0]), tot[2]))
print("%-22s %8s %6s %6s %6s  source" % ("line", "instr%", "lanes", "smp%", ""))
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f = os.path.join(ROOT, "pdsat_b200", "csrc", ln[0]) if ln and ln[0] != "?" else None
    text = ""
    if f and os.path.exists(f):
        if f not in src:
            src[f] = open(f).read().splitlines()
        if 0 < ln[1] <= len(src[f]):
            text = src[f][ln[1] - 1].strip()[:90]
    print("%-22s %7.2f%% %6.1f %5.1f%%         %s" % ("%s:%d" % ln, 100.0 * v[0] / tot[0], v[1] / max(1, v[0]), 100.0 * v[2] / max(1, tot[2]), text))

# optional: share of instruction count per line range "name:lo-hi,..." (argv[4])
if len(sys.argv) > 4:
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":")
        lo, hi = (int(x) for x in rng.split("-"))
        a = [0, 0, 0]
        for ln, v in agg.items():
            if ln and ln[0] == "kernels.cuh" and lo <= ln[1] <= hi:
                for j in range(3):
                    a[j] += v[j]
        print("%-14s lines %4d-%4d  instr %6.2f%%  lanes %5.1f  samples %5.1f%%" % (name, lo, hi, 100.0 * a[0] / tot[0], a[1] / max(1, a[0]), 100.0 * a[2] / max(1, tot[2])))
