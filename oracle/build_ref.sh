#!/usr/bin/env bash
# Build the UNMODIFIED reference hot path (hmc.c + mersenne_inline.c) from where it lies under
# /root/reference into oracle/_ref/ as shared objects the tests/bench can dlopen.  TEST INFRASTRUCTURE ONLY.
#
#   libhmcref_<NT>x<NX>_<flavour>[_nsN].so
#
# * No reference source is copied into the repo: the two unguarded lattice #defines (hmc.c:14-15) are
#   rewritten by sed in a pipe that feeds gcc on stdin.
# * flavour "compat"  = verbatim source (fm_conjugate_mul == fm_mul).
# * flavour "adjoint" = verbatim + the one-hunk correction inside fm_conjugate_mul only (hmc.c:197-248):
#   every hop sign flipped and exp(mu) <-> exp(-mu), i.e. the true M^dagger.
# * "-Dmain=hmc_main" turns the driver into a callable; -fPIC (semantic interposition left on) keeps every
#   internal call to fm_mul/fm_conjugate_mul/fmdm_invert_cg going through the PLT, so a library loaded
#   earlier can interpose them.
# * optional _nsN: 'int nsteps = 10;' (hmc.c:708) rewritten to N leapfrog steps.
set -euo pipefail
REF=${TB_REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OPT=${TB_REF_OPT:--O3 -march=x86-64-v3}
mkdir -p "$OUT"
if [ ! -f "$REF/hmc.c" ]; then
  echo "build_ref.sh: $REF/hmc.c not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi

build_one() { # NT NX flavour [nsteps] [shipped|-] [CG_MAX_ITER]
  local nt=$1 nx=$2 fl=$3 ns=${4:-10} opt=$OPT maxit=${6:-}
  local name="libhmcref_${nt}x${nx}_${fl}"
  [ -n "$maxit" ] && name="${name}_it${maxit}"
  [ "$ns" != 10 ] && name="${name}_ns${ns}"
  # the reference's own CFLAGS (Makefile:5 "-march=native -std=c99 -g", i.e. no optimisation); x86-64-v3 instead of
  # native so that the object also runs on the GPU box's host CPU
  if [ "${5:-}" = shipped ]; then name="${name}_shipped"; opt="-march=x86-64-v3 -g"; fi
  local sedprog="s/^#define NT 32/#define NT ${nt}/; s/^#define NX 32/#define NX ${nx}/; s/int nsteps = 10;/int nsteps = ${ns};/"
  # bounded CPU baselines at sizes where a full solve takes minutes to hours: "#define CG_MAX_ITER 100000" (hmc.c:35)
  # rewritten, so the reference's own fmdm_invert_cg runs N - 1 iterations and returns (max-iter is silent, hmc.c:364)
  [ -n "$maxit" ] && sedprog="$sedprog; s/^#define CG_MAX_ITER 100000/#define CG_MAX_ITER ${maxit}/"
  if [ "$fl" = adjoint ]; then
    sedprog="$sedprog; 197,248{s/v += 0\\.5/v @@ 0.5/; s/v -= 0\\.5/v += 0.5/; s/v @@ 0\\.5/v -= 0.5/; s/expmmu/EXPTMP/; s/expmu/expmmu/; s/EXPTMP/expmu/}"
  fi
  sed "$sedprog" "$REF/hmc.c" | gcc $opt -std=c99 -w -fPIC -shared -Dmain=hmc_main \
      -I"$HERE/shim" -I"$REF" -x c - -x none "$REF/mersenne_inline.c" "$HERE/shim/lapack_stub.c" \
      -o "$OUT/$name.so" -lm
}

SIZES=${TB_REF_SIZES:-"8x8 16x16 16x32 32x32 64x64 128x128 256x256"}
for s in $SIZES; do
  nt=${s%x*}; nx=${s#*x}
  build_one "$nt" "$nx" compat
  build_one "$nt" "$nx" adjoint
done
# light-mass trajectory oracle: 40 leapfrog steps
build_one 32 32 adjoint 40
build_one 64 64 adjoint 40
# CPU baseline at the flags the reference ships with
build_one 64 64 adjoint 10 shipped
# bounded CPU baselines of the large configurations: 40 iterations at 256^2, 2 at 2048^2 (bench.py other_configs)
build_one 256 256 adjoint 10 - 41
build_one 2048 2048 adjoint 10 - 3
# family B: vec_ops.c behind Thirring.h (sizes are unguarded #defines, Thirring.h:14-15).  The translation unit
# is assembled on gcc's stdin: "#define MAIN" (so that the EXTERN globals of Thirring.h:54-76 are DEFINED here,
# as the driver fermionbag.c does), the size-rewritten header, then vec_ops.c without its own #include.
build_vecops() { # NT NX [symmetric]: the boundary choice is a #define in Thirring.h:27-28
  local nt=$1 nx=$2 suffix="" bcsed=""
  if [ "${3:-}" = symmetric ]; then
    suffix="_symmetric"
    bcsed="; s|^#define ANTISYMMETRIC|//#define ANTISYMMETRIC|; s|^//#define SYMMETRIC|#define SYMMETRIC|"
  fi
  ( echo '#define MAIN'
    sed "s/^#define NT 64/#define NT ${nt}/; s/^#define NX 64/#define NX ${nx}/${bcsed}" "$REF/Thirring.h"
    sed '/#include "Thirring.h"/d' "$REF/vec_ops.c" ) | gcc $OPT -std=c99 -w -fPIC -shared \
      -I"$HERE/shim" -I"$REF" -x c - -x none "$REF/mersenne_inline.c" -o "$OUT/libvecopsref_${nt}x${nx}${suffix}.so" -lm
}
SIZES_B=${TB_REF_SIZES_B:-"16x16 32x32 64x64 16x32"}
for s in $SIZES_B; do build_vecops "${s%x*}" "${s#*x}"; done
for s in 16x32 64x64; do build_vecops "${s%x*}" "${s#*x}" symmetric; done
gcc -O2 -o "$OUT/ref_hmc" "$HERE/ref_launcher.c" -ldl
ls "$OUT" | sed 's/^/  built oracle\/_ref\//'
