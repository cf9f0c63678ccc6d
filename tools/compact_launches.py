#!/usr/bin/env python
"""Compacts an `ncu --metrics gpu__time_duration.sum --csv` log into id,kernel,grid,block,ns
plus a per-kernel summary: python tools/compact_launches.py in.csv out.csv "comment"."""
import collections
import csv
import re
import sys


def main():
    src, dst, comment = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ii, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "ID", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    out = []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        n = re.sub(r"^void ", "", r[ki])
        for ns in ("pdlp_b200::kernels::", "pdlp_b200::build_kernels::", "pdlp_b200::", "kernels::", "build_kernels::"):
            n = n.replace(ns, "")
        n = re.sub(r"\(.*$", "", n)[:80]
        t = float(r[vi].replace(",", ""))
        out.append((r[ii], n, r[gi], r[bi], t))
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n# per-kernel summary (count, total us, share):\n" % comment)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("#   %6d %12.1f %6.2f%%  %s\n" % (v[0], v[1] / 1e3, 100 * v[1] / total, k))
        f.write("id,kernel,grid,block,gpu_time_ns\n")
        for r in out:
            f.write('%s,"%s","%s","%s",%.0f\n' % r)


if __name__ == "__main__":
    main()
