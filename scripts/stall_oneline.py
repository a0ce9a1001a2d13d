#!/usr/bin/env python
"""stdin: one [stall] launch report (header + role lines) -> one line: label, plan, kernel clocks, busy % per role."""
import re
import sys

label = sys.argv[1]
lines = [l for l in sys.stdin.read().splitlines() if l.startswith("[stall]")]
if not lines:
    print(f"{label:9s} (no report)")
    sys.exit(0)
plan = lines[0].split("|", 1)[1].strip() if "|" in lines[0] else lines[0]
plan = re.sub(r"a1_st=\d+ acc1_st=\d+ acc2_st=\d+ |w_st=\d+ |grid=\d+ ctas=\d+", "", plan)
tot, roles = 0.0, []
for r in lines[1:]:
    if r.startswith("[stall] tc"):
        break
    name = r.split()[1]
    t = float(re.search(r"total\s+(\d+)", r).group(1))
    busy = float(re.search(r"busy\s+([\d.]+)%", r).group(1))
    tot = max(tot, t)
    roles.append(f"{name} {busy:.0f}%")
print(f"{label:9s} {tot / 1e3:8.1f} kclk | {plan} | " + "  ".join(roles))
