"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
    python scripts/launch_summary.py gpurun_out/launches.csv [out.json]"""
import collections
import csv
import json
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        v = float(r[hdr.index("Metric Value")].replace(",", ""))
        f = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[hdr.index("Metric Unit")], 1e-6)
        a = agg.setdefault(name, dict(launches=0, total_ms=0.0))
        a["launches"] += 1
        a["total_ms"] += v * f
    tot = sum(a["total_ms"] for a in agg.values())
    out = []
    for k, a in sorted(agg.items(), key=lambda x: -x[1]["total_ms"]):
        a.update(kernel=k, share=a["total_ms"] / tot, avg_ms=a["total_ms"] / a["launches"])
        out.append(a)
        print("%-32s launches=%4d total=%12.3f ms avg=%10.4f ms share=%6.2f%%" % (k, a["launches"], a["total_ms"], a["avg_ms"], 100 * a["share"]))
    if len(sys.argv) > 2:
        json.dump(dict(source=sys.argv[1], total_ms=tot, kernels=out), open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
