#!/bin/bash
# SASS opcode summary of the in-tree libpmb.so (evidence for which hardware paths the kernels use):
#   bash scripts/sass_summary.sh > profiles/sass_opcodes_r2.txt
so=pymoto_b200/libpmb.so
echo "# cuobjdump -sass $so  (sm_100a cubins), built $(date -u +%Y-%m-%dT%H:%MZ) from $(git rev-parse --short HEAD 2>/dev/null)"
cuobjdump -sass $so > /tmp/_sass.txt 2>/dev/null
echo "# arch lines: $(grep -c 'arch = sm_100a' /tmp/_sass.txt) x sm_100a"
echo "# opcode totals over all kernels (TMA: UBLKCP = 1-D bulk copies, UTMALDG = tensor-map loads; SYNCS = mbarrier ops; DMMA = FP64 tensor cores)"
grep -oE '^\s+/\*[0-9a-f]+\*/\s+[@!A-Z0-9_.]+(\s+[A-Z0-9_.]+)?' /tmp/_sass.txt | awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); print op}' | sort | uniq -c | sort -rn | awk '$2 ~ /^(UBLKCP|UTMALDG|UTMASTG|SYNCS|DMMA|DFMA|DADD|DMUL|LDGSTS|LDS|STS|LDG|STG|SHFL|BAR|MUFU|ATOMS|RED|HMMA|UTCHMMA|LDTM)$/ {printf "%10d %s\n",$1,$2}'
echo "# per kernel: TMA / mbarrier / DMMA / DFMA counts"
awk '/Function : /{name=$3} /UBLKCP/{a[name]++} /UTMALDG/{t[name]++} /SYNCS/{s[name]++} /DMMA/{d[name]++} /DFMA/{f[name]++} END{for (k in f) if (a[k]+t[k]+d[k]>0 || f[k]>300) printf "%6d UBLKCP %4d UTMALDG %5d SYNCS %5d DMMA %6d DFMA  %s\n", a[k],t[k],s[k],d[k],f[k],k}' /tmp/_sass.txt | sort -k10 | c++filt 2>/dev/null | cut -c1-200
