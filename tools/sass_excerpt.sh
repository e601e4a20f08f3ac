#!/bin/bash
# usage: tools/sass_excerpt.sh > profiles/r2_sass_mnemonics.txt   (here, no GPU needed)
# Per kernel of libsdb200.so: code size and how often the instructions that characterise the design occur in the SASS
# (warp votes / matches / reductions, 128-bit loads and stores, shared-memory traffic, FP64, integer divisions would show up
# as long MUFU.RCP + IMAD sequences, local-memory spills as LDL / STL).
cd "$(dirname "$0")/.."
lib=pysignalduino_b200/libsdb200.so
echo "cuobjdump -sass $lib (sm_100a), $(date -u +%Y-%m-%d)"
cuobjdump -sass $lib 2>/dev/null | awk '
/Function :/ { if (name != "") flush(); name=$3; n=0; delete c; next }
/^[ \t]+\/\*[0-9a-f]+\*\// { n++; ins=$2; sub(/;$/,"",ins); split(ins,a,"."); m=a[1];
  if (ins ~ /^@/) { ins=$3; split(ins,a,"."); m=a[1] }
  if (m=="VOTE"||m=="MATCH"||m=="REDUX"||m=="SHFL"||m=="LDL"||m=="STL"||m=="ATOMG"||m=="ATOMS"||m=="RED"||m=="POPC"||m=="FLO"||m=="SHF"||m=="LOP3"||m=="VIMNMX"||m=="VIMNMX3"||m=="DMUL"||m=="DFMA"||m=="DADD"||m=="F2I"||m=="I2F"||m=="MUFU"||m=="BAR"||m=="LDGSTS") c[m]++;
  if (ins ~ /^LDG.*128/) c["LDG.128"]++; else if (m=="LDG") c["LDG"]++;
  if (ins ~ /^STG.*128/) c["STG.128"]++; else if (m=="STG") c["STG"]++;
  if (ins ~ /^LDS.*128/) c["LDS.128"]++; else if (m=="LDS") c["LDS"]++;
  if (m=="STS") c["STS"]++;
  next }
function flush(   k, s) { s=""; for (k in c) s=s sprintf(" %s=%d", k, c[k]); printf "%-72s %6d instr (%5.1f KB) %s\n", name, n, n*16/1024, s }
END { flush() }' | sort
