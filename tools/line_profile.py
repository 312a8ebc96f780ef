"""Join an ncu --import-source report with nvdisasm line info: warp-instructions per source line (developer tool).
usage: line_profile.py <rep> <kernel-mangled-substring> <warp_diagonals>"""
import re, csv, io, collections, subprocess, sys, os, glob, tempfile
rep, ksub, diag = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "sedef_b200", "libsedef_b200.so")], cwd=tmp, capture_output=True)
dis, start = None, None
for cubin in sorted(glob.glob(os.path.join(tmp, "*sm_100a*.cubin"))):      # one cubin per translation unit: find the kernel's
    d = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    hit = [i for i, l in enumerate(d) if l.startswith(".text.") and ksub in l]
    if hit:
        dis, start = d, hit[0]
        break
assert dis is not None, "kernel not found in any cubin of libsedef_b200.so"
cur, insts = None, []
for l in dis[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and insts:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        insts.append((m2.group(2), cur))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
assert len(data) == len(insts), (len(data), len(insts))
agg, sagg, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for k, (txt, f) in enumerate(insts):
    key = f or ("?", 0)
    v = float(data[k][ia]) / diag
    agg[key] += v; sagg[key] += float(data[k][isamp])
    op = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt.strip()).group(2)
    ops[key][op] += v
text = {fn: open(os.path.join(root, "sedef_b200", "csrc", fn)).read().split("\n") for fn in ("extz_dp.cuh", "extz_core.cuh", "extz_dp16.cuh")}
tot, ts = sum(agg.values()), sum(sagg.values()) or 1
print("total warp-instructions per warp-diagonal: %.1f" % tot)
for key, v in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 70):
    fn, ln = key
    t = text.get(fn, [""] * 1)[ln - 1].strip()[:80] if fn in text and ln > 0 else ""
    top = ",".join("%s:%.1f" % (o, c) for o, c in ops[key].most_common(3))
    print("%6.1f %5.1f%%  %s:%d  %-80s | %s" % (v, 100 * sagg[key] / ts, fn, ln, t, top))
