#!/usr/bin/env python
"""Condense a FV_STALL_DEBUG report (gpurun_out/stall_*.txt): one line per launch with per-role clk/unit and busy %."""
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
i = 0
while i < len(lines):
    l = lines[i]
    if not l.startswith("[stall] tc"):
        i += 1
        continue
    m = re.match(r"\[stall\] (tc\d) (.*?) \| (.*?) ctas=(\d+)", l)
    kind, shape, plan, ctas = m.groups()
    shape = re.sub(r"Lpos=\d+ |L=\d+ |B=\d+ |layout=", "", shape)
    roles = []
    j = i + 1
    while j < len(lines) and not lines[j].startswith("[stall] tc"):
        r = lines[j]
        name = r.split()[1]
        tot = float(re.search(r"total\s+(\d+)", r).group(1))
        units = float(re.search(r"units\s+([\d.]+)", r).group(1))
        busy = float(re.search(r"busy\s+([\d.]+)%", r).group(1))
        waits = re.findall(r"(\w+)\s+([\d.]+)%", r)
        top = max((w for w in waits if w[0] != "busy"), key=lambda w: float(w[1]), default=("", "0"))
        roles.append(f"{name}:{busy:.0f}%busy({tot / max(units, 1e-9):.0f}/u) wait {top[0]} {float(top[1]):.0f}%")
        tot_clk = tot
        j += 1
    print(f"{kind} {shape} | {plan}")
    print("     " + " | ".join(roles))
    i = j
