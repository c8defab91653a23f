#!/bin/bash
# TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference (yuansliu/minicom) from
# the sources where they lie under /root/reference/src into oracle/_ref/, one
# binary per (read length, mode), with link-time wrappers (mcref_wrap.cpp) for
# phase timing and state dumps.  Nothing is copied from the reference; the only
# generated file is config.h, which the reference's own wrapper script also
# generates per run (minicom:56-91, :187-213).
#
# usage: build_ref.sh <readlen> <sg|order|pe>
# output: oracle/_ref/minicom_ref_L<readlen>_<mode>   (+ build_L<..>/ objects)
#         oracle/_ref/decompress                      (once)
set -euo pipefail
L=$1; MODE=$2
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${MC_REFERENCE_SRC:-/root/reference/src}
OUT=$HERE/../_ref
B=$OUT/build_L${L}_${MODE}
[ -d "$REF" ] || { echo "reference sources not found at $REF" >&2; exit 3; }
mkdir -p "$B"
{
  echo "#pragma once"
  echo "#include \"mcb_cfg.h\""
  if [ "$MODE" = order ]; then echo "#define ORDER"; echo "int cmpcluster3(const void *a_, const void *b_);"; fi
  if [ "$MODE" = pe ]; then echo "#define _PE"; echo "int cmpcluster3(const void *a_, const void *b_);"; fi
  echo "#define readlen $L"
  echo "#define num_thr mcb_cfg_int(\"MC_T\", 1)"
  echo "#define uniqid mcb_cfg_str(\"MC_UNIQID\", \"umc\")"
  echo "#define output mcb_cfg_str(\"MC_TMPDIR\", \"output_mc/\")"
  echo "#define inik mcb_cfg_int(\"MC_K\", 0)"
  echo "#define inithr mcb_cfg_int(\"MC_E\", 0)"
  echo "#define inimaxthr mcb_cfg_int(\"MC_EMAX\", 0)"
  echo "#define inistep mcb_cfg_int(\"MC_STEP\", 0)"
  echo "#define ininumdict mcb_cfg_int(\"MC_S\", 0)"
  echo "#define iniw mcb_cfg_int(\"MC_W\", 0)"
  echo "#define inim mcb_cfg_int(\"MC_M\", 0)"
  echo "#define inicbthr mcb_cfg_int(\"MC_CBTHR\", 0)"
  echo "#define inimaxrounds mcb_cfg_int(\"MC_MAXROUNDS\", 0)"
} > "$B/config.h"
# -march=x86-64-v3 instead of the reference's -march=native so the binary also runs on the GPU box's host CPU
CXXFLAGS="-O3 -std=c++11 -w -march=x86-64-v3 -fopenmp -I$B -I$HERE -I$HERE/../../dropin -I$REF"
OBJS="bseq misc preprocess sketch bbhashdict kthread_reads kthread_bucket kthread_idx kthread_cb kthread_dump kthread_hash_realign minicommain"
# the reference links minicompe from an archive (src/Makefile:24-25,33-34): kthread_dump.o is never pulled in there and
# defines the same cmp() as kthread_dump_pe.o
[ "$MODE" = pe ] && OBJS="${OBJS/kthread_dump /kthread_dump_pe }"
pids=()
for f in $OBJS; do g++ $CXXFLAGS -c "$REF/$f.c" -o "$B/$f.o" & pids+=($!); done
g++ $CXXFLAGS -c "$HERE/mcref_wrap.cpp" -o "$B/mcref_wrap.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
WRAP="-Wl,--wrap=_Z12kt_for_readsiP7reads_tl -Wl,--wrap=_Z13kt_for_bucketiP7reads_tl -Wl,--wrap=_Z17mm_idx_generationiP8mm_idx_t -Wl,--wrap=_Z15combine_clusteriP7reads_tPi -Wl,--wrap=_Z12realign_hashiP7reads_tii"
ALL=""; for f in $OBJS; do ALL="$ALL $B/$f.o"; done
g++ -O3 -fopenmp $ALL "$B/mcref_wrap.o" $WRAP -o "$OUT/minicom_ref_L${L}_${MODE}" -lm -lz -lpthread
if [ ! -x "$OUT/decompress" ]; then
  g++ -O3 -std=c++11 -w -march=x86-64-v3 -fopenmp -include stdint.h -I$B -I$HERE -I$HERE/../../dropin -I$REF "$REF/decompress.c" -o "$OUT/decompress" -lm -lz -lpthread
fi
echo "built $OUT/minicom_ref_L${L}_${MODE}"
# CPU check of the dump writer (dropin/mcb_dump_writer.h, N2): the reference with its kt_dump_[pe_]for weakened and replaced by the
# writer + the oracle's print_encode (mcref_dumpcheck.cpp).  usage: build_ref.sh <readlen> <mode> dumpcheck
if [ "${3:-}" = dumpcheck ]; then
  DUMP=kthread_dump; SYM=_Z11kt_dump_foriP7reads_ti
  [ "$MODE" = pe ] && { DUMP=kthread_dump_pe; SYM=_Z14kt_dump_pe_foriP7reads_ti; }
  NCALL=$(objdump -dr "$B/$DUMP.o" | grep -c "R_X86_64_PLT32[[:space:]]*$SYM" || true)
  [ "$NCALL" -ge 1 ] || { echo "$DUMP.o does not call $SYM through its symbol" >&2; exit 4; }
  objcopy --weaken-symbol=$SYM "$B/$DUMP.o" "$B/${DUMP}_weak.o"
  g++ $CXXFLAGS -c "$HERE/mcref_dumpcheck.cpp" -o "$B/mcref_dumpcheck.o"
  gcc -O2 -std=c99 -c "$HERE/../mc_oracle.c" -o "$B/mc_oracle.o"
  ALLW="${ALL/$B\/$DUMP.o/$B\/${DUMP}_weak.o}"
  g++ -O3 -fopenmp $ALLW "$B/mcref_wrap.o" "$B/mcref_dumpcheck.o" "$B/mc_oracle.o" $WRAP -o "$OUT/minicom_ref_L${L}_${MODE}_dumpcheck" -lm -lz -lpthread
  echo "built $OUT/minicom_ref_L${L}_${MODE}_dumpcheck"
fi
# unit-level entry points of the reference (hash64, mm_sketch_two, mm_sketch_lh_ori, radix_sort_128x, bseq_read) for ctypes tests
if [ ! -f "$OUT/libmcref_units.so" ] || [ "$HERE/mcref_units.cpp" -nt "$OUT/libmcref_units.so" ] || [ "$HERE/mcref_units_bseq.cpp" -nt "$OUT/libmcref_units.so" ]; then
  g++ -O2 -std=c++11 -w -fPIC -shared -I$B -I$HERE -I$HERE/../../dropin -I$REF "$HERE/mcref_units.cpp" "$HERE/mcref_units_sort.cpp" "$HERE/mcref_units_bseq.cpp" "$REF/misc.c" "$REF/bseq.c" -o "$OUT/libmcref_units.so" -lz
  echo "built $OUT/libmcref_units.so"
fi
