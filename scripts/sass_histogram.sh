#!/bin/bash
# Driver-visible proof of the compiled state: SASS mnemonic histogram of the shipped library (tcgen05 = UTCHMMA /
# UTCQMMA..., TMEM = LDTM / STTM, TMA = UTMALDG / UTMASTG / UTMAREDG) and ptxas register / spill lines per kernel.
# Usage: scripts/sass_histogram.sh r02   -> profiles/r02_sass_histogram.txt, profiles/r02_ptxas.txt
R=${1:-r02}
cd "$(dirname "$0")/.."
LIB=vtamiq_b200/libvtamiq_b200.so
{
  echo "# cuobjdump -sass $LIB | mnemonic histogram ($(date -u +%F), nvcc 12.9)"
  echo "## Blackwell-specific mnemonics"
  cuobjdump -sass $LIB | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z][A-Z0-9_.]+" | awk '{print $NF}' \
    | grep -E "^(UTC|LDTM|STTM|UTMA|UBLKCP|SYNCS|UTCBAR|MUFU|HMMA|FFMA2|FADD2|FMUL2|F2FP|FMNMX3|ACQBULK|USETMAXREG)" | sort | uniq -c | sort -rn
  echo "## all mnemonics (base opcode)"
  cuobjdump -sass $LIB | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z][A-Z0-9_]+" | awk '{print $NF}' | sort | uniq -c | sort -rn | head -60
} > profiles/${R}_sass_histogram.txt
{
  echo "# ptxas -v per kernel (vtamiq_b200/ptxas.log, built by vtamiq_b200/build.py)"
  grep -E "Compiling entry|Used [0-9]+ registers|spill" vtamiq_b200/ptxas.log | sed -E "s/ptxas info    : //" \
    | paste - - - | sed -E "s/Compiling entry function '([^']+)' for 'sm_100a'/\1/" | while read -r line; do
      name=$(echo "$line" | awk '{print $1}' | c++filt); echo "$name | $(echo "$line" | cut -f2- )"; done
} > profiles/${R}_ptxas.txt
wc -l profiles/${R}_sass_histogram.txt profiles/${R}_ptxas.txt
