#!/bin/sh
# TEST INFRASTRUCTURE — builds oracle/_ref/ribbit_ref_cp: the reference with
# CP1/CP2 checkpoint logging. The reference sources are copied
# to a throw-away temp dir, the top-level call sites of the three
# addSeedToSeedPositions* functions (those AFTER the scan functions start, so
# recursive calls stay unlogged) are redirected to the RB_CP1_* macros of
# cp_hooks.h, and a CP2 dump is added before the per-seed stage. Only the
# binary is written into the repo tree (git-ignored oracle/_ref/). CP4: every call of mostFrequentLongerMotif
# (parse_seed.cpp:153) logs its arguments and the row it chose (mmotif_index, parse_seed.cpp:241).
set -e
OUTBIN="$1"
: "${REF:=/root/reference}" "${CXX:=g++}" "${CC:=gcc}"
T=$(mktemp -d)
trap 'rm -rf "$T"' EXIT
cp "$REF"/*.cpp "$REF"/*.h "$REF"/ssw.c "$T"/
# scan functions start at these lines (parse_perfect_shiftxor.cpp:146, parse_substitute_shiftxor.cpp:391,
# parse_anchored_shiftxor.cpp:538); verify the anchors before patching
sed -n '146p' "$T/parse_perfect_shiftxor.cpp"    | grep -q 'processShiftXORsPerfect'           || { echo "anchor P moved"; exit 1; }
sed -n '391p' "$T/parse_substitute_shiftxor.cpp" | grep -q 'processShiftXORswithSubstitutions' || { echo "anchor S moved"; exit 1; }
sed -n '538p' "$T/parse_anchored_shiftxor.cpp"   | grep -q 'processShiftXORsAnchored'          || { echo "anchor A moved"; exit 1; }
sed -n '74p'  "$T/fasta_utils.cpp"               | grep -q 'START_TIME = time(0)'              || { echo "anchor F moved"; exit 1; }
sed -n '170p' "$T/fasta_utils.cpp"               | grep -q 'considering indels'                || { echo "anchor F2 moved"; exit 1; }
sed -n '245p' "$T/parse_seed.cpp"                | grep -q 'return mmotif_index'               || { echo "anchor M moved"; exit 1; }
sed -i '245a rb_cp_rec(21, seed_start, seed_sequence_length, motif_length, mmotif_index);' "$T/parse_seed.cpp"
sed -i '147,$ s/addSeedToSeedPositionsPerfect(/RB_CP1_P(/'        "$T/parse_perfect_shiftxor.cpp"
sed -i '392,$ s/addSeedToSeedPositionsSubstitutions(/RB_CP1_S(/'  "$T/parse_substitute_shiftxor.cpp"
sed -i '539,$ s/addSeedToSeedPositionsAnchored(/RB_CP1_A(/'       "$T/parse_anchored_shiftxor.cpp"
sed -i '170a rb_cp2_dump(seed_positions_perfect, seed_positions_substut, seed_positions_anchored); if (getenv("RB_CP_STOP_AFTER_CP2")) return;' "$T/fasta_utils.cpp"
sed -i '74a rb_cp_contig_start((long)sequence.length());' "$T/fasta_utils.cpp"
$CC -O3 -w -c "$T/ssw.c" -o "$T/ssw.o"
for f in ssw_cpp global_variables process_cigar parse_seed parse_smallmotif_seed merge_types \
         parse_anchored_shiftxor parse_substitute_shiftxor parse_perfect_shiftxor bitseq_utils fasta_utils ribbit; do
    $CXX -O3 -w -I"$SHIM" -I"$T" -include "$HOOKS" -c "$T/$f.cpp" -o "$T/$f.o"
done
$CXX -O3 "$T"/*.o -o "$OUTBIN"
