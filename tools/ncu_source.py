"""Rank the SASS of one kernel from `ncu --page source --csv --print-source sass` by executed instructions / stall samples.

usage: ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_source.py src.csv [top]
"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    tot_i = sum(int(d["Instructions Executed"]) for d in data)
    tot_s = sum(int(d["# Samples"]) for d in data)
    print(f"instructions executed {tot_i:,}  stall samples {tot_s:,}  sass lines {len(data)}")
    by_op = collections.Counter()
    by_op_s = collections.Counter()
    for d in data:
        src = d["Source"].strip()
        src = re.sub(r"^@!?U?P\d+\s+", "", src)
        op = src.split()[0].split(".")[0] if src else "?"
        by_op[op] += int(d["Instructions Executed"])
        by_op_s[op] += int(d["# Samples"])
    print("-- by opcode (inst share, sample share)")
    for op, n in by_op.most_common(top):
        print(f"  {op:12s} {100*n/tot_i:6.2f}%  {100*by_op_s[op]/max(tot_s,1):6.2f}%")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for d in data:
        for h in stalls:
            agg[h] += int(d[h] or 0)
    print("-- stall reasons (all samples)")
    for h, n in agg.most_common(10):
        print(f"  {h:26s} {100*n/max(tot_s,1):6.2f}%")
    print("-- hottest SASS lines by samples")
    for d in sorted(data, key=lambda d: -int(d["# Samples"]))[:top]:
        print(f"  {int(d['# Samples']):7d} {int(d['Instructions Executed']):10d}  {d['Source'].strip()[:90]}")


if __name__ == "__main__":
    main()
