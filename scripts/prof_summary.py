#!/usr/bin/env python
"""Summarise a bench.py --profile-out JSON by layer shape."""
import collections, json, sys
d = json.load(open(sys.argv[1]))
agg = collections.OrderedDict()
for r in d["layers"]:
    k = (r["kernel"], r["Cin"], r["N"], r["K"], r["dil"])
    a = agg.setdefault(k, [0, 0.0, 0.0, 0, 0.0])
    a[0] += 1; a[1] += r["ms"]; a[2] += r["flops"]; a[3] = r["positions"]; a[4] += r["bytes"]
tot = sum(a[1] for a in agg.values())
print("total ms %.2f" % tot)
for k, a in agg.items():
    print("%-8s Cin=%3d N=%4d K=%2d d=%d n=%2d ms=%7.3f (%4.1f%%) algTF=%6.1f GB/s(unfused)=%6.0f ns/pos=%.3f" % (
        k[0], k[1], k[2], k[3], k[4], a[0], a[1], 100 * a[1] / tot, a[2] / a[1] / 1e9, a[4] / a[1] / 1e6, a[1] / a[0] * 1e6 / a[3]))
