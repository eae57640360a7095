#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference (mrottmann/DDalphaAMG) from where it
# lies under /root/reference into oracle/_ref/libddref.so (+ libddref_sse.so), against the single-rank MPI shim in
# oracle/ref_shim/.  Recipe follows the reference's own Makefile (Makefile:25-29,82-98): every *_generic.{c,h} is
# instantiated with float.sed / double.sed, the rest is taken as is; flags -std=gnu99 -O3 -ffast-math -fopenmp.
# Generated sources live only in a temporary directory; nothing from the reference is copied into this repo.
# The reference's own build system is not run.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${DDA_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF not present (GPU box uses the prebuilt oracle/_ref)"; exit 0; fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/ddref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/gsrc"
for f in "$REF"/src/*; do
  b="$(basename "$f")"
  case "$b" in
    *_generic.c|*_generic.h)
      sed -f "$REF/float.sed"  "$f" > "$TMP/gsrc/${b/_generic/_float}"
      sed -f "$REF/double.sed" "$f" > "$TMP/gsrc/${b/_generic/_double}" ;;
    *) cp "$f" "$TMP/gsrc/$b" ;;
  esac
done
cp "$HERE/ref_harness.c" "$TMP/gsrc/zz_ref_harness.c"
cp "$HERE/ref_shim/mpi_shim.c" "$TMP/gsrc/zz_mpi_shim.c"
COMMON="-std=gnu99 -O3 -ffast-math -fopenmp -DOPENMP -DPARAMOUTPUT -DPROFILING -fPIC -w -I$HERE/ref_shim -I$TMP/gsrc"
build_flavour () {  # $1 = name, $2 = extra flags
  local name="$1" extra="$2" od="$TMP/obj_$1"
  mkdir -p "$od"
  # main.c has its own main()+globals, dd_alpha_amg.c is included textually by the harness
  ls "$TMP"/gsrc/*.c | grep -v -E '/(main|dd_alpha_amg)\.c$' | \
    xargs -P "$(nproc)" -I{} sh -c 'gcc '"$COMMON $extra"' -c "$1" -o "'"$od"'/$(basename "$1" .c).o"' _ {}
  gcc -shared -fopenmp -o "$OUT/$name" "$od"/*.o -lm
  echo "built $OUT/$name"
}
build_flavour libddref.so ""
if [ "${DDA_REF_SSE:-1}" = "1" ]; then build_flavour libddref_sse.so "-DSSE -msse4.2" || echo "SSE flavour failed (non-fatal)"; fi
