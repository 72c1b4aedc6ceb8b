#!/bin/bash
# SASS census of libffb200.so (what proves the Blackwell-native paths, B200_PROFILING.md): usage scratch/sass_census.sh > profiles/rNN_sass_census.md
so=factor-fields_b200/libffb200.so
cuobjdump -sass $so > /tmp/ffb_sass.txt 2>/dev/null
echo "# SASS census of \`$so\` (commit $(git rev-parse --short HEAD))"
echo
echo "\`cuobjdump -sass $so | grep -c <mnemonic>\`; cubin architectures: $(cuobjdump -lelf $so | grep -o 'sm_[0-9a-z]*' | sort | uniq -c | tr '\n' ' ')"
echo
echo "| SASS mnemonic | count | written as |"
echo "|---|---:|---|"
for row in "UTCHMMA|tcgen05.mma.kind::f16 (tensor cores, TMEM accumulators)" "LDTM|tcgen05.ld (TMEM -> registers)" "UBLKCP|cp.async.bulk (TMA engine, no tensor map)" "UTMALDG|cp.async.bulk.tensor (tensor-map TMA loads)" "SYNCS|mbarrier operations" "REDG.E.ADD.F32x4|red.global.add.v4.f32" "REDG.E.ADD.F32x2|red.global.add.v2.f32" "VOTE|warp ballots (compaction, run detection)" "SHFL|warp shuffles (scans)" " HMMA|legacy mma.sync (must be 0)"; do
  m="${row%%|*}"; d="${row#*|}"; echo "| \`$m\` | $(grep -c -- "$m" /tmp/ffb_sass.txt) | $d |"
done
echo
echo "Kernels containing UTCHMMA:"
awk '/Function :/{f=$3} /UTCHMMA/{c[f]++} END{for(k in c) print "- `" k "`: " c[k]}' /tmp/ffb_sass.txt | sort
