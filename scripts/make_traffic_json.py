"""profiles/traffic.json from ncu CSV captures of the bench launch sizes (dram__bytes_read.sum + dram__bytes_write.sum per
launch of the walk kernel and of the skip-gram kernel):

    python scripts/make_traffic_json.py tract24=profiles/r2s12_traffic_tract24.csv synth100k=profiles/r2s12_traffic_synth100k.csv

The captures come from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
-k regex:'k_sgns|k_walk_alias' --csv --log-file <csv> python bench.py [--workload W] --steps 1 --warmup 0 --no-e2e --no-synth --no-cpu-baseline`."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = {}
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    for arg in sys.argv[1:]:
        wl, fn = arg.split("=")
        rows = [r for r in csv.reader(open(os.path.join(ROOT, fn))) if len(r) > 14 and r[0].isdigit()]
        per = collections.defaultdict(dict)
        for r in rows:
            per[(r[0], r[4])][r[12]] = float(r[14])
        for (_, k), m in per.items():
            name = "k_walk_alias" if "k_walk_alias" in k else "sgns"
            b = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
            e = out.setdefault(wl, {})
            if name not in e or b > e[name]["bytes"]:   # the largest launch per kernel (the flow corpus, not the spatial one)
                e[name] = dict(bytes=int(b), read=int(m.get("dram__bytes_read.sum", 0)), write=int(m.get("dram__bytes_write.sum", 0)),
                               time_ns=m.get("gpu__time_duration.sum"), l2_hit_pct=m.get("lts__t_sector_hit_rate.pct"), kernel=k, source=fn, commit=commit)
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
