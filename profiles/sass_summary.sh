#!/bin/bash
# Blackwell-native evidence: counts of tcgen05 / TMEM / TMA SASS mnemonics per kernel of the in-tree library.
#   bash profiles/sass_summary.sh > profiles/r2_sass_summary.txt
SO=${1:-pert_gnn_kdd23_b200/libpertgnn.so}
echo "# cuobjdump -sass $SO ($(date -u +%FT%TZ)); columns: UTCHMMA (tcgen05.mma) | LDTM+STTM (tcgen05.ld/st) | UTMALDG (TMA tensor load) | UTMASTG (TMA tensor store) | UTMAREDG (TMA reduce-add store) | UBLKCP (1-D bulk copy) | SYNCS (mbarrier) | REDG/RED (global reductions) | instructions"
cuobjdump -sass "$SO" | awk '
/Function :/ { if (name != "") print_row(); name=$3; mma=ldst=tld=tst=tred=blk=syn=red=ins=0; next }
/^[ \t]*\/\*[0-9a-f]+\*\// { ins++ }
/UTCHMMA/ { mma++ } /LDTM|STTM/ { ldst++ } /UTMALDG/ { tld++ } /UTMASTG/ { tst++ } /UTMAREDG/ { tred++ } /UBLKCP/ { blk++ } /SYNCS/ { syn++ } /REDG|RED\./ { red++ }
function print_row() { if (mma+ldst+tld+tst+tred+blk > 0 || name ~ /k_tile_|k_segreduce_stream|k_allreduce|k_store_nodes/) printf "%-110s %4d %4d %3d %3d %3d %3d %4d %4d %6d\n", name, mma, ldst, tld, tst, tred, blk, syn, red, ins }
END { print_row() }' | c++filt | sed 's/(anonymous namespace):://; s/_GLOBAL__N__[0-9a-f_]*cu_[0-9a-f]*//'
