#!/usr/bin/env bash
# Build the UNMODIFIED reference solver sources (read-only at /root/reference) into
# shared libraries under oracle/_ref/ so that tests and bench.py can call the
# reference's own sim_solv()/solve_*()/tree_build() through ctypes.
#
# TEST INFRASTRUCTURE ONLY.  Nothing in the product path links or loads these.
#
# The reference is configured exclusively through #defines in src/config.h (no
# #ifndef guards, quote-includes), so each variant is produced by copying src/ to
# a scratch directory OUTSIDE the repo (mktemp under /tmp), sed-editing config.h
# there, and compiling with gcc.  Only the resulting .so lands in the repo tree
# (oracle/_ref/ is git-ignored but travels to the GPU box).  No reference source
# is copied into the repo.
#
# Variants:
#   pp        ENABLE_TREE off                      -> solve_gravity_part_part
#   tree      ENABLE_TREE on, TREE_NODES_MAX=2^20  -> tree_build + solve_gravity_part_tree
#   tree_big  ENABLE_TREE on, TREE_NODES_MAX=2^23  (N up to 2^24)
#   pp_sph    ENABLE_TREE off, ENABLE_SPH on       -> + solve_sph
#   tree_vis  like tree with SOLVER_VIS 3 left ON  -> solve_gravity_part_tree(sim, render_buffer, sizex, sizey)
#             draws its debug overlay (gravity.c:16-111,380-445,550-574).  The only symbol that needs is matMVP, which
#             lives in visualization.c (not compiled here): a one-line definition is generated in the scratch directory.
#   vis       like tree, plus the reference's visualization.c (vis_init / vis_draw / vis_free): the
#             particle splat of vis_draw() with the shipped SPATIAL_SAMPLES 128, EXPOSURE, ORTHO_SCALE, AUTOSCALING and
#             the tanh tonemap (VIS_FILE_FORMAT 2); the I/O thread and the OpenMP target pragma are off.
# In all other variants: ENABLE_VIS / SOLVER_VIS / VIS_FILENAME off (no rasteriser dependency),
# CSV logs and timing prints off, Kahan flags ON (shipped default), fp32.
#
# Flags: the reference Makefile's performance set is -O3 -march=native -flto=auto -fopenmp -std=c99
# (Makefile:27-28,58,88).  -march=native is replaced by x86-64-v3 because the .so is
# built in the dev container and executed on the GPU box's (different) host CPU; the
# hot loops are scalar (volatile Kahan) so the ISA level does not change their speed,
# and -std=c99 keeps -ffp-contract=off so results are bit-identical.
set -euo pipefail
REF=${SF_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
CC=${SF_REF_CC:-/usr/bin/gcc}
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF/src not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
SCRATCH="$(mktemp -d /tmp/sfref.XXXXXX)"
trap 'rm -rf "$SCRATCH"' EXIT

build_variant() {
  local name="$1" tree="$2" sph="$3" nodes="$4" vis="${5:-0}"
  local d="$SCRATCH/$name"
  mkdir -p "$d"
  cp "$REF"/src/*.c "$REF"/src/*.h "$d"/
  local cfg="$d/config.h"
  sed -i -E \
    -e 's|^#define ENABLE_VIS$|//#define ENABLE_VIS|' \
    -e 's|^#define ENABLE_OFFLOAD_VIS$|//#define ENABLE_OFFLOAD_VIS|' \
    -e 's|^#define ENABLE_VIS_IO_THREAD$|//#define ENABLE_VIS_IO_THREAD|' \
    -e 's|^#define SOLVER_VIS 3$|//#define SOLVER_VIS 3|' \
    -e 's|^#define VIS_FILENAME |//#define VIS_FILENAME |' \
    -e 's|^#define ENABLE_TIMING$|//#define ENABLE_TIMING|' \
    -e 's|^#define LOG_STATISTICS |//#define LOG_STATISTICS |' \
    -e 's|^#define LOG_TIMINGS_SIM_STEP |//#define LOG_TIMINGS_SIM_STEP |' \
    -e 's|^#define LOG_TIMINGS_VIS_DRAW |//#define LOG_TIMINGS_VIS_DRAW |' \
    -e 's|^#define LOG_TIMINGS_FFT_EXEC |//#define LOG_TIMINGS_FFT_EXEC |' \
    -e "s|^#define TREE_NODES_MAX .*|#define TREE_NODES_MAX $nodes|" \
    "$cfg"
  if [ "$tree" = 0 ]; then sed -i -E 's|^#define ENABLE_TREE$|//#define ENABLE_TREE|' "$cfg"; fi
  if [ "$sph" = 1 ]; then sed -i -E 's|^//#define ENABLE_SPH$|#define ENABLE_SPH|' "$cfg"; fi
  local extra=""
  if [ "$vis" = 1 ]; then
    sed -i -E 's|^//#define SOLVER_VIS 3$|#define SOLVER_VIS 3|' "$cfg"
    printf '#include "types.h"\nreal matMVP[16];\n' > "$d/sfref_matmvp.c"
    extra="sfref_matmvp.c"
  fi
  if [ "$vis" = 2 ]; then extra="visualization.c"; fi
  ( cd "$d" && "$CC" -shared -fPIC -fopenmp -O3 -march=x86-64-v3 -flto=auto -std=c99 -w \
      simulation.c gravity.c tree.c sph.c initcond.c log.c grid.c $extra \
      -o "$OUT/libsfref_$name.so" -lm )
  echo "built $OUT/libsfref_$name.so"
}

build_variant pp       0 0 262144
build_variant tree     1 0 1048576
build_variant tree_big 1 0 8388608
build_variant pp_sph   0 1 262144
build_variant tree_vis 1 0 1048576 1
build_variant vis      1 0 1048576 2
