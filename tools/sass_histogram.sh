# SASS opcode histogram of the product library (no GPU needed): bash tools/sass_histogram.sh > profiles/sass_opcodes.txt
LIB=pointnet_refine_b200/liblrn_b200.so
echo "# SASS opcode histogram of $LIB (cuobjdump -sass, sm_100a), commit $(git rev-parse --short HEAD)+"
echo "# tcgen05 = UTCHMMA* (MMA), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit), UTCATOMSWS (TMEM alloc); TMA = UTMALDG / UTMASTG; mbarrier = SYNCS.*; HMMA = mma.sync (train attention)"
echo "# count  opcode"
cuobjdump -sass $LIB > /tmp/lrn_sass.txt
ops() { grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/;$//'; }
ops < /tmp/lrn_sass.txt | sort | uniq -c | sort -rn
echo
echo "# per kernel: tcgen05 / TMA / mma.sync instructions"
awk '/Function :/ {name=$3} /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// {print name, $0}' /tmp/lrn_sass.txt | \
  grep -E "UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|HMMA" | \
  sed -E 's/^(\S+)\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?(\S+).*/\1 \3/' | sed 's/;$//' | sort | uniq -c | \
  while read n k op; do echo "$n $op $(echo $k | c++filt | cut -c1-90)"; done | sort -k3,3 -k1,1nr
