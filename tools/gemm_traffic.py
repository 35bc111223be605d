"""Aggregate an ncu metrics pass over the gemm_bf16_kernel launches of `bench.py --profile-run` into
profiles/gemm_traffic_r2.json (the `roofline.traffic` source of bench.py) and a per-shape table.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none -k regex:gemm_bf16_kernel --csv --log-file gpurun_out/gemm_traffic.csv python bench.py --profile-run
    python tools/gemm_traffic.py gpurun_out/gemm_traffic.csv <launches_per_step> [out.json] [table.txt]

The JSON is keyed on the hash of the kernel / op sources (bench.source_sha16), the workload string and the batch: bench.py
quotes the figure only for the very build and workload it was captured on.
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    path, per_step = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "gemm_traffic_r2.json")
    table = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "gemm_traffic_r2_by_shape.txt")
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    by_id = collections.OrderedDict()
    for r in rows:
        if "gemm_bf16_kernel" not in r["Kernel Name"]:
            continue
        d = by_id.setdefault(r["ID"], {"name": r["Kernel Name"].split("gemm_bf16_kernel")[1].split("(")[0], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d[r["Metric Name"] + ":unit"] = r["Metric Unit"]
    launches = list(by_id.values())[-per_step:]           # the second (timed) step of --profile-run
    assert len(launches) == per_step, (len(launches), per_step)

    def to_bytes(d, k):
        v, u = d.get(k, 0.0), d.get(k + ":unit", "byte").lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    def to_ms(d):
        v, u = d.get("gpu__time_duration.sum", 0.0), d.get("gpu__time_duration.sum:unit", "ns").lower()
        return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(u, 1e-6)

    rd = sum(to_bytes(d, "dram__bytes_read.sum") for d in launches)
    wr = sum(to_bytes(d, "dram__bytes_write.sum") for d in launches)
    ms = sum(to_ms(d) for d in launches)
    tkey = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tens = sum(d.get(tkey, 0.0) * to_ms(d) for d in launches) / max(ms, 1e-9)

    import bench
    import argparse
    a = argparse.Namespace(layers=0, speakers=int(os.environ.get("SPEAKERS", "2")), seconds=float(os.environ.get("SECONDS_", "10")),
                           mode=os.environ.get("MODE", "train"))
    js = {"launches_per_step": per_step, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
          "gpu_time_ms_under_ncu": ms, "tensor_pipe_active_pct_time_weighted": tens,
          "source_sha16": bench.source_sha16(), "workload": bench.workload_name(a), "batch": int(os.environ.get("BATCH", "32")),
          "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum," + tkey +
                    " --clock-control none -k regex:gemm_bf16_kernel python bench.py --profile-run (second step)"}
    with open(out, "w") as f:
        json.dump(js, f, indent=1)
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for d in launches:
        key = (d["name"], d["grid"], round(to_ms(d) * 1e3, -1))
        a_ = agg[key]
        a_[0] += 1
        a_[1] += to_ms(d)
        a_[2] += to_bytes(d, "dram__bytes_read.sum")
        a_[3] += to_bytes(d, "dram__bytes_write.sum")
        a_[4] += d.get(tkey, 0.0) * to_ms(d)
    with open(table, "w") as f:
        f.write(f"gemm_bf16_kernel, one step: {per_step} launches, {ms:.2f} ms under ncu, DRAM read {rd / 1e9:.1f} GB write {wr / 1e9:.1f} GB, "
                f"tensor pipe active (time-weighted) {tens:.1f} %\n")
        f.write(f"{'kernel<CFG,NCTA>':>18s} {'grid':>12s} {'~us':>7s} {'n':>4s} {'ms':>8s} {'rd GB':>8s} {'wr GB':>8s} {'TB/s':>6s} {'tensor%':>8s}\n")
        for (name, grid, us), (n, t, r, w, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:>18s} {grid:>12s} {us:7.0f} {n:4d} {t:8.3f} {r / 1e9:8.2f} {w / 1e9:8.2f} {(r + w) / 1e9 / max(t, 1e-9):6.2f} {tp / max(t, 1e-9):8.1f}\n")
    print(open(table).read())
    print(json.dumps(js))


if __name__ == "__main__":
    main()
