"""Per-instruction stall summary of an .ncu-rep captured with --set full --import-source on:
    python scripts/ncu_stalls.py rep.ncu-rep [top_n]
Prints the stall-reason shares over all samples, the top instructions by samples with their dominant reason, and the
executed opcode mix."""
import collections
import csv
import subprocess
import sys


def main():
    rep, top_n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print(rows[0][1] if len(rows[0]) > 1 else rows[0])
    hdr, data = rows[1], rows[2:]
    g = hdr.index
    isrc, isamp, iex = g("Source"), g("# Samples"), g("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda r, i: int(r[i] or 0)
    tot = sum(num(r, isamp) for r in data)
    print("samples %d, SASS instructions %d, executed warp instructions %d" % (tot, len(data), sum(num(r, iex) for r in data)))
    agg = {s: sum(num(r, g(s)) for r in data) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
        print("  %-26s %5.1f %%" % (s, 100.0 * v / tot))
    print("top instructions by samples:")
    for r in sorted(data, key=lambda r: -num(r, isamp))[:top_n]:
        best = max(stalls, key=lambda s: num(r, g(s)))
        print("  %5.2f %%  executed %11d  %-22s %s" % (100.0 * num(r, isamp) / tot, num(r, iex), best, " ".join(r[isrc].split())[:90]))
    mix = collections.Counter()
    for r in data:
        op = r[isrc].split()
        if op:
            mix[(op[1] if op[0].startswith("@") else op[0]).split(".")[0]] += num(r, iex)
    te = sum(mix.values())
    print("opcode mix (executed):", ", ".join("%s %.1f %%" % (o, 100.0 * v / te) for o, v in mix.most_common(12)))


if __name__ == "__main__":
    main()
