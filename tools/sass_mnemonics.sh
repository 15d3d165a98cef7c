#!/bin/bash
# per-kernel instruction counts and the Blackwell-specific mnemonics of the built library -> profiles/sass_mnemonics_<tag>.txt
tag=${1:-r02}
cuobjdump -sass vine_robot_isaacgymenvs_b200/csrc/libvine_b200.so > /tmp/lib.sass
python - "$tag" <<'PY'
import re, subprocess, sys
from collections import Counter
txt = open('/tmp/lib.sass').read()
rows = []
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0].strip()
    ops = re.findall(r'/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', f)
    c = Counter(ops)
    keys = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "UTMALDG", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "LDL", "STL"]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r'\((?:[^()]|\([^()]*\))*\)\s*$', '', dem).replace("(anonymous namespace)::", "").replace("void ", "")
    rows.append((dem, len(ops), {k: c[k] for k in keys if c[k]}))
with open(f'profiles/sass_mnemonics_{sys.argv[1]}.txt', 'w') as out:
    out.write("# cuobjdump -sass csrc/libvine_b200.so (sm_100a): instructions per kernel and counts of the mnemonics that matter\n")
    out.write("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA), UTCBAR = tcgen05.commit,\n"
              "# FFMA2/FMUL2/FADD2 = packed fma/mul/add.rn.f32x2; regenerate: tools/sass_mnemonics.sh\n")
    for name, n, c in sorted(rows):
        out.write(f"{name:44s} {n:6d} instr  " + " ".join(f"{k}={v}" for k, v in c.items()) + "\n")
PY
