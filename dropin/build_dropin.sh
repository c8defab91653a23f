#!/bin/bash
# Builds the drop-in executable: the reference's KEPT host objects, compiled from the sources where they lie under
# /root/reference/src (nothing is copied), linked with dropin/mcb_dropin.cpp and libminicom_b200.so in place of
# sketch.o kthread_reads.o kthread_bucket.o kthread_idx.o kthread_hash_realign.o bbhashdict.o.
#
# usage: build_dropin.sh <readlen> <sg|order|pe> [wrap]
#   output: dropin/_build/minicom_b200_L<readlen>_<mode>
#   with `wrap`: additionally links oracle/ref/mcref_wrap.cpp (state dumps for the parity tests) ->
#           oracle/_ref/minicom_b200_L<readlen>_<mode>_wrapped      (test infrastructure)
set -euo pipefail
L=$1; MODE=$2; WRAPPED=${3:-}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/.." && pwd)
REF=${MC_REFERENCE_SRC:-/root/reference/src}
[ -d "$REF" ] || { echo "reference sources not found at $REF" >&2; exit 3; }
B=$HERE/_build/L${L}_${MODE}
mkdir -p "$B"
{
  echo "#pragma once"
  echo "#include \"mcb_cfg.h\""
  if [ "$MODE" = order ]; then echo "#define ORDER"; echo "int cmpcluster3(const void *a_, const void *b_);"; fi
  if [ "$MODE" = pe ]; then echo "#define _PE"; echo "int cmpcluster3(const void *a_, const void *b_);"; fi
  echo "#define readlen $L"
  for kv in "num_thr MC_T 1" "inik MC_K 0" "inithr MC_E 0" "inimaxthr MC_EMAX 0" "inistep MC_STEP 0" "ininumdict MC_S 0" "iniw MC_W 0" "inim MC_M 0" "inicbthr MC_CBTHR 0" "inimaxrounds MC_MAXROUNDS 0"; do
    set -- $kv; echo "#define $1 mcb_cfg_int(\"$2\", $3)"
  done
  echo "#define uniqid mcb_cfg_str(\"MC_UNIQID\", \"umc\")"
  echo "#define output mcb_cfg_str(\"MC_TMPDIR\", \"output_mc/\")"
} > "$B/config.h"
CXXFLAGS="-O3 -std=c++11 -w -march=x86-64-v3 -fopenmp -I$B -I$HERE -I$REF -I$ROOT/include"
# bseq.o (the FASTQ reader) is replaced by the shim's packing reader (N3) unless MCB_KEEP_BSEQ=1
KEPT="misc preprocess kthread_cb kthread_dump minicommain"
if [ "${MCB_KEEP_BSEQ:-0}" = 1 ]; then KEPT="bseq $KEPT"; CXXFLAGS="$CXXFLAGS -DMCB_KEEP_BSEQ"; fi
if [ "${MCB_KEEP_DUMP:-0}" = 1 ]; then CXXFLAGS="$CXXFLAGS -DMCB_KEEP_DUMP"; fi
# minicompe links from an archive (src/Makefile:24-25,33-34): kthread_dump.o is never pulled in and clashes with kthread_dump_pe.o
[ "$MODE" = pe ] && KEPT="${KEPT/kthread_dump /kthread_dump_pe }"
pids=()
for f in $KEPT; do g++ $CXXFLAGS -c "$REF/$f.c" -o "$B/$f.o" & pids+=($!); done
g++ $CXXFLAGS -c "$HERE/mcb_dropin.cpp" -o "$B/mcb_dropin.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
# kthread_cb.o stays linked (cmpcluster2/3, match_pro and construct_ref2 are used by the kept dump stage), but its combine_cluster
# is renamed so that the shim's combine_cluster (the GPU contig merge) is the one preprocess.o calls; the reference's own host
# merge remains reachable as mcb_ref_combine_cluster (MCB_HOST_MERGE=1, and the multi-GPU path).
objcopy --redefine-sym _Z15combine_clusteriP7reads_tPi=_Z23mcb_ref_combine_clusteriP7reads_tPi "$B/kthread_cb.o"
# N2: kthread_dump[_pe].o stays linked for cluster_dump[_pe] (info.txt, the single-read and id files), but its per-contig worker
# kt_dump_[pe_]for is WEAKENED so that the shim's definition (device encoding + dropin/mcb_dump_writer.h) is the one cluster_dump
# calls; a second copy of the object with everything else made local keeps the reference's worker reachable as
# mcb_ref_kt_dump_[pe_]for (MCB_HOST_DUMP=1).  MCB_KEEP_DUMP=1 at build time leaves the object alone.
EXTRA=""
if [ "${MCB_KEEP_DUMP:-0}" != 1 ]; then
  DUMP=kthread_dump; SYM=_Z11kt_dump_foriP7reads_ti; NEW=_Z19mcb_ref_kt_dump_foriP7reads_ti
  [ "$MODE" = pe ] && { DUMP=kthread_dump_pe; SYM=_Z14kt_dump_pe_foriP7reads_ti; NEW=_Z22mcb_ref_kt_dump_pe_foriP7reads_ti; }
  NCALL=$(objdump -dr "$B/$DUMP.o" | grep -c "R_X86_64_PLT32[[:space:]]*$SYM" || true)
  [ "$NCALL" -ge 1 ] || { echo "$DUMP.o does not call $SYM through its symbol: cannot interpose the dump worker" >&2; exit 4; }
  # (only the object's own strong definitions are made local: COMDAT / weak items such as DW.ref.__gxx_personality_v0 must stay
  # as they are, pthread_exit in the reference's workers unwinds through them)
  LARGS=""
  for s in $(nm "$B/$DUMP.o" | awk '$2 ~ /^[TBDR]$/ {print $3}' | grep -v "^$SYM\$"); do LARGS="$LARGS --localize-symbol=$s"; done
  objcopy --redefine-sym $SYM=$NEW $LARGS "$B/$DUMP.o" "$B/${DUMP}_refworker.o"
  objcopy --weaken-symbol=$SYM "$B/$DUMP.o"
  EXTRA="$B/${DUMP}_refworker.o"
fi
ALL=""; for f in $KEPT; do ALL="$ALL $B/$f.o"; done
LIBDIR=$ROOT/minicom_b200
g++ -O3 -fopenmp $ALL $EXTRA "$B/mcb_dropin.o" -L"$LIBDIR" -lminicom_b200 -Wl,-rpath,'$ORIGIN/../../minicom_b200' -o "$HERE/_build/minicom_b200_L${L}_${MODE}" -lm -lz -lpthread
echo "built $HERE/_build/minicom_b200_L${L}_${MODE}"
if [ "$WRAPPED" = wrap ]; then
  g++ $CXXFLAGS -c "$ROOT/oracle/ref/mcref_wrap.cpp" -o "$B/mcref_wrap.o"
  WRAP="-Wl,--wrap=_Z12kt_for_readsiP7reads_tl -Wl,--wrap=_Z13kt_for_bucketiP7reads_tl -Wl,--wrap=_Z17mm_idx_generationiP8mm_idx_t -Wl,--wrap=_Z15combine_clusteriP7reads_tPi -Wl,--wrap=_Z12realign_hashiP7reads_tii"
  mkdir -p "$ROOT/oracle/_ref"
  g++ -O3 -fopenmp $ALL $EXTRA "$B/mcb_dropin.o" "$B/mcref_wrap.o" $WRAP -L"$LIBDIR" -lminicom_b200 -Wl,-rpath,'$ORIGIN/../../minicom_b200' -o "$ROOT/oracle/_ref/minicom_b200_L${L}_${MODE}_wrapped" -lm -lz -lpthread
  echo "built $ROOT/oracle/_ref/minicom_b200_L${L}_${MODE}_wrapped"
fi
