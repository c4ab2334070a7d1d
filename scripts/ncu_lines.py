"""Per-source-line instruction and stall-sample shares of an .ncu-rep (needs -lineinfo and --import-source on):
    python scripts/ncu_lines.py rep.ncu-rep [top_n] [file-substring]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    lines = []
    fpath = ""
    for r in rows:
        if r and r[0] == "File Path":
            fpath = r[1]
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            inst = int(r[hdr.index("Instructions Executed")])
            samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        lines.append((fpath.split("/")[-1], int(r[0]), r[1].strip(), inst, samp))
    ti, ts = sum(x[3] for x in lines) or 1, sum(x[4] for x in lines) or 1
    print("total warp instructions %d, samples %d" % (ti, ts))
    for f, ln, src, inst, samp in sorted(lines, key=lambda x: -x[3])[:top]:
        print("%-12s %4d  inst %5.2f%%  samples %5.2f%%  %s" % (f, ln, 100.0 * inst / ti, 100.0 * samp / ts, src[:110]))


if __name__ == "__main__":
    main()
