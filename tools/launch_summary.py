"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --profile-run` per kernel (last step)."""
import csv, collections, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if 'conv0_kernel' in x['Kernel Name']]
st = rows[idx[-1]:]
tot = sum(float(x['Metric Value']) for x in st)
agg = collections.defaultdict(lambda: [0, 0.0])
for x in st:
    n = re.sub(r'\(.*', '', x['Kernel Name']).replace('void ', '')
    if 'at::native' in n:
        n = 'torch:' + n.split('at::native::')[-1][:60]
    if 'gemm_bf16_kernel' in n:
        n = 'mtasr::gemm_bf16_kernel'
    agg[n][0] += 1
    agg[n][1] += float(x['Metric Value'])
print(f"timed step: {len(st)} launches, {tot / 1e6:.3f} ms of kernel time")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}% {c:5d}  {n[:100]}")
