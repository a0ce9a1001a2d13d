#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library (cuobjdump -sass; no GPU needed):
profiles/<pre>_sass_<kernel>.txt = mnemonic histogram + every tensor-core / TMEM / TMA / bulk-copy / mbarrier instruction with
its address; profiles/<pre>_sass_<kernel>.sass.gz = the complete listing.

    python scripts/sass_listing.py r02
"""
import collections
import gzip
import os
import re
import subprocess
import sys

pre = sys.argv[1] if len(sys.argv) > 1 else "r02"
LIB = "fastvocoder_b200/_C/libfastvocoder_b200.so"
KEY = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|SYNCS|UTCCP|REDG|RED\.|ELECT|ACQBULK|NANOSLEEP)")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, cur, name = {}, None, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        cur = kernels.setdefault(name, [])
        continue
    if cur is not None:
        cur.append(line)
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WANT = ("conv_tc3_fused_kernel", "conv_tc2_kernel", "conv_narrow7", "pqmf_synthesis_poly", "pqmf_analysis_v4", "pack_split", "encode16")
index = []
for mangled, lines in kernels.items():
    dn = demangle(mangled)
    if not any(w in dn for w in WANT) or "<true" in dn or "<(bool)1" in dn:   # skip the FV_STALL_DEBUG instantiations
        continue
    short = re.sub(r"[^A-Za-z0-9_]+", "_", dn.split("(")[0].replace("void ", "").replace("fv::", ""))[:70].strip("_")
    ins = [l for l in lines if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
    ops = collections.Counter()
    keyl = []
    for l in ins:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
        if not m:
            continue
        ops[m.group(2).split(".")[0]] += 1
        if KEY.search(l):
            keyl.append(re.sub(r"\s+", " ", l.split("/*", 1)[1].replace("*/", "", 1)).split("/*")[0].strip())
    path = os.path.join("profiles", f"{pre}_sass_{short}.txt")
    with open(path, "w") as f:
        f.write(f"{dn}\n{len(ins)} SASS instructions (cuobjdump -sass {LIB}, sm_100a)\n\nmnemonic histogram:\n")
        for k, v in ops.most_common():
            f.write(f"  {k:14s} {v}\n")
        f.write(f"\ntensor-core / TMEM / TMA / bulk-copy / mbarrier instructions ({len(keyl)}):\n")
        for l in keyl:
            f.write("  " + l + "\n")
    with gzip.open(os.path.join("profiles", f"{pre}_sass_{short}.sass.gz"), "wt") as f:
        f.write("\n".join(lines))
    c = lambda k: sum(v for kk, v in ops.items() if kk.startswith(k))
    index.append(f"{short:72s} {len(ins):6d} instr  UTCHMMA {c('UTCHMMA'):4d}  LDTM {c('LDTM'):3d}  UTMALDG {c('UTMALDG'):3d}  UBLKCP {c('UBLKCP'):3d}  UTCBAR {c('UTCBAR'):3d}  SYNCS {c('SYNCS'):4d}")
open(os.path.join("profiles", f"{pre}_sass_index.txt"), "w").write("\n".join(sorted(index)) + "\n")
print("\n".join(sorted(index)))
