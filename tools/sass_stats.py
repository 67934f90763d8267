#!/usr/bin/env python
"""Per-kernel SASS statistics of the built library (instruction mix, longest runs of global loads)."""
import re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "manifold_mcmc_for_diffusions_b200", "libmmd_b200.so")
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None; ops = {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1); ops[fn] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn: ops[fn].append(m.group(1))
for fn, o in ops.items():
    if pat and pat not in fn: continue
    if not o: continue
    runs = []; cur = 0
    for x in o:
        if x.startswith("LDG"): cur += 1
        else:
            if cur: runs.append(cur)
            cur = 0
    cnt = lambda p: sum(x.startswith(p) for x in o)
    print(f"{fn[:70]:70s} n={len(o):6d} LDG={cnt('LDG'):4d} maxrun={sorted(runs, reverse=True)[:6]} STG={cnt('STG'):4d} DFMA={cnt('DFMA'):5d} DMUL={cnt('DMUL'):4d} DADD={cnt('DADD'):4d} LDL={cnt('LDL'):4d} STL={cnt('STL'):4d} MUFU={cnt('MUFU'):3d} BAR={cnt('BAR'):3d}")
