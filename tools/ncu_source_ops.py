"""Per-opcode instruction and stall-sample totals from `ncu -i X.ncu-rep --page source --csv`."""
import collections
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    si, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, stall = collections.Counter(), collections.Counter()
    lines = []
    for r in rows[h + 1:]:
        try:
            sass, n, st = r[si], int(r[ie]), int(r[ws])
        except Exception:
            continue
        toks = sass.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += n
        stall[op.split(".")[0]] += st
        lines.append((st, n, sass))
    tot, ts = sum(ops.values()), max(sum(stall.values()), 1)
    print("total warp-instructions %d, stall samples %d" % (tot, ts))
    for op, n in ops.most_common(top):
        print("%-12s %10d %5.1f%%   stalls %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * stall[op] / ts))
    print("--- top stall lines")
    for st, n, sass in sorted(lines, reverse=True)[:25]:
        print("%6d %9d  %s" % (st, n, sass[:110]))


if __name__ == "__main__":
    main(sys.argv[1])
