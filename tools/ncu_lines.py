"""Aggregate an ncu --set full report per CUDA source line (warp-stall samples), using nvdisasm line info of the object.
usage: python tools/ncu_lines.py report.ncu-rep object.o kernel_substring [top]"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
callsite = len(sys.argv) > 5 and sys.argv[5] == "callsite"
main_file = re.sub(r"(_v\d+)?\.o$", ".cu", os.path.basename(obj))
last_main = None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
# several profiled launches in one report: keep the first section only
for i, r in enumerate(data):
    if r and r[0] == "Kernel Name":
        data = data[:i]
        break
data = [r for r in data if len(r) == len(hdr)]
ix = {k: i for i, k in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur_fn, cur_line, infn = None, None, False
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur_line:
        # inlined helpers (ptx.cuh, common.cuh, CUDA headers) are charged to the last line of the kernel's own file
        if cur_line[0] == main_file:
            last_main = cur_line
        line_of[int(m.group(1), 16)] = cur_line if (cur_line[0] == main_file or last_main is None or not callsite) else (last_main[0], last_main[1])
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    ln = line_of.get(off, ("?", 0))
    n = int(r[ix["# Samples"]])
    tot += n
    a = agg[ln]
    a[0] += n
    a[1] += int(r[ix["Instructions Executed"]])
    for k in stalls:
        v = int(r[ix[k]])
        if v:
            a[2][k] += v
src_cache = {}
def src(fn, n):
    for d in ("multi-talker-asr-with-llms_b200/csrc", "."):
        p = os.path.join(d, fn)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
    return ""
print("total samples", tot, "mapped lines", len(agg))
for ln, (n, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    main = ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(2))
    print(f"{n:6d} {100 * n / tot:5.1f}%  x{ins:>9}  {ln[0]}:{ln[1]:<4d} {src(*ln):90s} [{main}]")
