#!/bin/bash
# rebuild the library and summarise the pass kernels' SASS (registers, spills, size, opcode mix)
cd "$(dirname "$0")/.."
python fastecc_b200/build.py --force 2>&1 | grep -E "error|spill|Compiling.*ntt_pass" | paste - - | sed 's/ptxas info    ://g; s/Compiling entry function//; s/for .sm_100a.//' | grep -E "Li10ELi1|Li9ELi2|error"
cuobjdump -sass fastecc_b200/libfastecc_b200.so > /tmp/all.sass
for pat in "ILi10ELi1ELi1" "ILi9ELi2ELi0"; do
  s=$(grep -n "Function : .*ntt_pass_kernel${pat}" /tmp/all.sass | cut -d: -f1)
  e=$(awk -v s=$s 'NR>s && /Function :/{print NR; exit}' /tmp/all.sass); [ -z "$e" ] && e=$(wc -l < /tmp/all.sass)
  sed -n "${s},${e}p" /tmp/all.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's/\/\*[0-9a-f]*\*\/\s*$//' | awk '{ $1=""; print}' | sed 's/\/\*.*//' > /tmp/k_${pat}.sass
  f=/tmp/k_${pat}.sass
  echo "$pat instrs: $(wc -l < $f)  LDL/STL: $(grep -c 'LDL\|STL' $f)  IMAD.HI: $(grep -c IMAD.HI $f)  BAR: $(grep -c BAR.SYNC $f) MOVs: $(grep -c -E '^ *(MOV|IMAD.MOV|HFMA2)' $f)"
done
