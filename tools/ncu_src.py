"""Summarise an ncu --set full report: stall reasons, hottest SASS lines, instruction mix (run where ncu is installed)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print("kernel:", rows[0][1][:100]); print("total samples", tot, "SASS lines", len(data))
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(r[ix[k]]) for r in data) for k in stalls}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:28s} {v:7d} {100 * v / max(tot, 1):5.1f}%")
print("hottest lines:")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
    st = {k: int(r[ix[k]]) for k in stalls if int(r[ix[k]]) > 0}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {r[ix['# Samples']]:>6} x{r[ix['Instructions Executed']]:>9}  {r[ix['Source']].strip()[:64]:64s} {main}")
mix = collections.Counter()
for r in data:
    src = r[ix['Source']].strip().split()
    if not src: continue
    op = src[1] if src[0].startswith('@') and len(src) > 1 else src[0]
    mix[op.split('.')[0]] += int(r[ix['Instructions Executed']])
tt = sum(mix.values())
print("warp instructions", tt, " mix:", ", ".join(f"{k} {100 * v / tt:.1f}%" for k, v in mix.most_common(14)))
