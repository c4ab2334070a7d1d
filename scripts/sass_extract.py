"""Opcode histogram of the shipped hot kernels from the built libdge.so (cuobjdump -sass; runs on the CPU box), so that
the instruction-level claims of DESIGN.md (one LDG.256 per walk step, 128-bit REDG row updates, peer loads / stores of
the exchange kernel, no tensor-core or TMA opcode anywhere: neither stage is a contraction or a tile copy) are
reviewable without rebuilding.  Writes profiles/<tag>_sass_opcodes.txt.

    python scripts/sass_extract.py [tag]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "embedding_b200", "libdge.so")
KERNELS = ["k_walk_alias", "k_walk_cdf", "k_sgns_sentILi8ELb0ELi0E", "k_sgns_sentILi32ELb0ELi0E", "k_sgns_blockILi8ELb0ELi192E", "k_sgns_items_v2ILi8ELb0ELb0E", "k_sgns_items_v2ILi32ELb0ELb0E", "k_sgns_items_tpILi2E",
           "k_sgns_items_g4ILi1ELb0E", "k_sgns_seqILi1ELi5E", "k_dp_exchange_peerILi1E", "k_alias_small", "k_seq_format"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[cur].append(line)
    out = ["libdge.so SASS (sm_100a), %d kernels; opcode histograms of the hot ones (cuobjdump -sass)\n" % len(funcs)]
    allops = collections.Counter()
    for name, lines in funcs.items():
        for ln in lines:
            op = re.sub(r"^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?", "", ln).split()[0].rstrip(";")
            allops[op.split(".")[0]] += 1
    out.append("every kernel together, opcodes by family: " + ", ".join("%s %d" % kv for kv in allops.most_common(40)))
    tensor = [op for op in allops if re.match(r"(UTC.*MMA|HMMA|HGMMA|QGMMA|IGMMA|LDTM|STTM|UTMALDG|UTMASTG)", op)]
    out.append("tensor-core / TMEM / TMA-tensor opcodes present: %s (none expected: no dense contraction, no tile copy on this path)\n" % (tensor or "none"))
    for key in KERNELS:
        for name, lines in funcs.items():
            if key not in name:
                continue
            ops = collections.Counter()
            full = collections.Counter()
            for ln in lines:
                body = re.sub(r"^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?", "", ln)
                op = body.split()[0].rstrip(";")
                ops[op.split(".")[0]] += 1
                if re.match(r"(LDG|STG|REDG|ATOMG|LDGSTS|LDS|STS|SHFL|UBLK|LD\b|ST\b|RED|LDC|MUFU|DFMA|DMUL|DADD)", op):
                    full[op] += 1
            out.append("== %s  (%d instructions)" % (name, len(lines)))
            out.append("   families: " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))
            out.append("   memory / special opcodes: " + ", ".join("%s x%d" % kv for kv in sorted(full.items(), key=lambda kv: -kv[1])))
    path = os.path.join(ROOT, "profiles", "%s_sass_opcodes.txt" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
