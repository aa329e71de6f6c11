#!/bin/bash
# rebuild the library and summarise the pass kernel's SASS (registers, spills, size, opcode mix)
cd /root/repo
python fastecc_b200/build.py --force 2>&1 | grep -E "error|registers|spill" | head -3
cuobjdump -sass fastecc_b200/libfastecc_b200.so > /tmp/all.sass
e=$(grep -n "Function :" /tmp/all.sass | sed -n 2p | cut -d: -f1)
sed -n "37,${e}p" /tmp/all.sass | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's/\/\*[0-9a-f]*\*\/\s*$//' | awk '{ $1=""; print}' | sed 's/\/\*.*//' > /tmp/k.sass
echo "instrs: $(wc -l < /tmp/k.sass)  LDL/STL: $(grep -c 'LDL\|STL' /tmp/k.sass)  IMAD.HI: $(grep -c IMAD.HI /tmp/k.sass)  BAR: $(grep -c BAR.SYNC /tmp/k.sass)"
