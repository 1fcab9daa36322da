#!/bin/bash
# Census of the SASS of the built library: which memory / synchronisation instructions the kernels really use
# (256-bit streaming loads, bulk async copies + mbarriers, cluster barriers, warp match) and that no tensor-core
# instruction is present (the path is byte/bit work bound by HBM -- DESIGN.md section 3).
#     bash tools/sass_census.sh > profiles/r02_sass_census.txt
set -eu
LIB=${1:-mamri_pose_estimation_b200/libmamri_b200.so}
T=$(mktemp)
cuobjdump -sass "$LIB" > "$T"
echo "# $LIB: $(grep -c 'Function :' "$T") kernels, $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' "$T") SASS instructions (sm_100a)"
echo "# tensor-core instructions (HMMA|IMMA|QMMA|UTCMMA|UTCHMMA|UTCQMMA): $(grep -cE 'HMMA|IMMA|QMMA|UTCMMA|UTCHMMA|UTCQMMA' "$T" || true)"
echo "# instruction counts by mnemonic (memory, atomics, synchronisation, bit work)"
grep -oE "\b(LDG|STG|LDS|STS|ATOMG|ATOMS|RED|REDG|UBLKCP|SYNCS|MATCH|REDUX|VOTE|VOTEU|SHFL|POPC|FLO|BREV|LOP3|CCTL|ERRBAR|MEMBAR|ACQBULK|UCGABAR_ARV|UCGABAR_WAIT|BAR|DFMA|DADD|DMUL)[.A-Z0-9_]*" "$T" | sort | uniq -c | sort -rn
echo "# kernels holding LDG.E.NA.ENL2.256 (256-bit streaming loads with an L2 evict-first policy):"
awk '/Function :/ {f=$3} /LDG.E.NA.ENL2.256/ {c[f]++} END {for (k in c) print c[k], k}' "$T" | sort -rn | c++filt | cut -c1-160
echo "# kernels holding UBLKCP (cp.async.bulk) + SYNCS (mbarrier):"
awk '/Function :/ {f=$3} /UBLKCP/ {c[f]++} END {for (k in c) print c[k], k}' "$T" | sort -rn | c++filt | cut -c1-160
rm -f "$T"
