#!/usr/bin/env bash
# Build the UNMODIFIED reference driver (main.c, simulation.c, simulation_io.c, initcond.c, log.c, grid.c,
# visualization*.c) twice from /root/reference/src:
#   _build/starflood_<variant>_ref   with its own gravity.c tree.c sph.c          (pure reference)
#   _build/starflood_<variant>_b200  with the three shims + libsfengine.so instead (drop-in)
# so that the two can be run side by side on the same box and their RAW snapshots compared
# (tests/test_gpu_dropin.py).  Reference sources are copied to a scratch dir under /tmp only; config.h is
# sed-edited there (the reference has no run-time configuration).
#
# Variant tree_vis keeps the reference's default ENABLE_VIS + SOLVER_VIS 3 (config.h:29,32): every step vis_draw()
# renders the particles, sim_step() adds the tree walk's debug overlay of particle #3 (gravity.c:380-445,550-574; in the
# drop-in: sfe_bh_overlay) and main.c saves the frame - as PFM (float RGB) here, so that frames compare bit for bit.  In the
# drop-in vis_draw() itself is vis_shim.c -> sfe_vis_draw (the particle splat with the shipped SPATIAL_SAMPLES 128).
#
# usage: build_dropin.sh <pp|tree|pp_sph|tree_vis> <N> <steps> [TREE_NODES_MAX]
set -euo pipefail
VARIANT=${1:?variant}; N=${2:?N}; STEPS=${3:?steps}; NODES=${4:-1048576}
REF=${SF_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/_build"
CC=/usr/bin/gcc
[ -d "$REF/src" ] || { echo "build_dropin.sh: $REF/src not present; keeping prebuilt binaries" >&2; exit 0; }
mkdir -p "$OUT"
S="$(mktemp -d /tmp/sfdropin.XXXXXX)"; trap 'rm -rf "$S"' EXIT
cp "$REF"/src/*.c "$REF"/src/*.h "$S"/
sed -i -E \
  -e 's|^#define ENABLE_VIS$|//#define ENABLE_VIS|' \
  -e 's|^#define ENABLE_VIS_IO_THREAD$|//#define ENABLE_VIS_IO_THREAD|' \
  -e 's|^#define SOLVER_VIS 3$|//#define SOLVER_VIS 3|' \
  -e 's|^#define VIS_FILENAME |//#define VIS_FILENAME |' \
  -e 's|^#define ENABLE_STB_IMAGE_WRITE$|//#define ENABLE_STB_IMAGE_WRITE|' \
  -e 's|^#define VIS_FILE_FORMAT 2$|#define VIS_FILE_FORMAT 0|' \
  -e "s|^#define NUM_BODIES .*|#define NUM_BODIES $N|" \
  -e "s|^#define NUM_TIMESTEPS .*|#define NUM_TIMESTEPS $STEPS|" \
  -e "s|^#define TREE_NODES_MAX .*|#define TREE_NODES_MAX $NODES|" \
  "$S/config.h"
case "$VARIANT" in
  pp)     sed -i -E 's|^#define ENABLE_TREE$|//#define ENABLE_TREE|' "$S/config.h" ;;
  tree)   ;;
  tree_vis) sed -i -E -e 's|^//#define ENABLE_VIS$|#define ENABLE_VIS|' -e 's|^//#define SOLVER_VIS 3$|#define SOLVER_VIS 3|' \
                -e 's|^//#define VIS_FILENAME |#define VIS_FILENAME |' -e 's|^#define ENABLE_OFFLOAD_VIS$|//#define ENABLE_OFFLOAD_VIS|' "$S/config.h" ;;
  pp_sph) sed -i -E -e 's|^#define ENABLE_TREE$|//#define ENABLE_TREE|' -e 's|^//#define ENABLE_SPH$|#define ENABLE_SPH|' "$S/config.h" ;;
  *) echo "unknown variant $VARIANT" >&2; exit 1 ;;
esac
FLAGS="-fopenmp -O3 -march=x86-64-v3 -std=c99 -w"
HOST="main.c simulation.c simulation_io.c initcond.c log.c grid.c visualization.c visualization_io.c"
( cd "$S" && $CC $FLAGS $HOST gravity.c tree.c sph.c -o "$OUT/starflood_${VARIANT}_ref" -lm -lpthread )
# drop-in: visualization.c keeps vis_init() / vis_free() / the matrix globals, its vis_draw() is renamed out of the way and
# vis_shim.c provides the one main.c calls (-> sfe_vis_draw)
HOST_B200="main.c simulation.c simulation_io.c initcond.c log.c grid.c visualization_io.c"
( cd "$S" && $CC $FLAGS -Dvis_draw=vis_draw_reference -c visualization.c -o visualization_novisdraw.o )
( cd "$S" && $CC $FLAGS -I"$S" -I"$ROOT/include" -I"$HERE" $HOST_B200 visualization_novisdraw.o "$HERE/gravity_shim.c" "$HERE/tree_shim.c" "$HERE/sph_shim.c" "$HERE/vis_shim.c" \
    -o "$OUT/starflood_${VARIANT}_b200" -L"$ROOT/starflood_b200" -lsfengine -Wl,-rpath,'$ORIGIN/../..' -lm -lpthread )
echo "built $OUT/starflood_${VARIANT}_ref and $OUT/starflood_${VARIANT}_b200 (N=$N, steps=$STEPS)"
