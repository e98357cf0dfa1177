#!/usr/bin/env bash
# Build the REFERENCE's own in-tree rasterizer (my_ext/_C/src/nerf/gaussian_*.cu, unmodified, compiled from
# where the sources lie under /root/reference) into oracle/_ref/ as torch/pybind11 modules:
#   _ref_raster        nvcc defaults (a*b+c contracted into FMAs) - what a user of the reference runs; the speed
#                      baseline of `bench.py --impl reference` and the tolerance-level parity target
#   _ref_raster_nofma  the SAME unmodified sources with -fmad=false: every product and sum separately rounded, i.e. the
#                      reference's arithmetic as its source text states it.  This is the build radii / keys / point_list /
#                      tile ranges are compared with BIT-EXACTLY (tests/test_gpu_reference_ext.py): nvcc's choice of
#                      which a*b+c to contract is a compiler artefact no second implementation can reproduce.
# The reference's CMake build uses CUDA separable compilation (computeColorFromSH is a cross-TU __device__
# function), hence -rdc=true.
# This is TEST/BENCH INFRASTRUCTURE: a second, GPU-side parity reference and the "reference extension" speed
# baseline.  No reference source is copied into this repo; only the compiled .so lands in
# oracle/_ref/ (git-ignored, travels to the GPU box).  We do NOT run the reference's own build system
# (its CMake route needs Eigen3, absent) - the hot-path files only need the vendored glm + torch headers.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/my_ext/_C/src/nerf" ]; then
  echo "[build_ref] $REF not present - keeping prebuilt $OUT (if any)"; exit 0
fi
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_LIB PY_INC EXT_SUFFIX CXX11_ABI <<<"$($PY - <<'PYEOF'
import sysconfig, torch
from torch.utils import cpp_extension as ce
inc = " ".join("-I" + p for p in ce.include_paths())
print(inc.replace(" ", ";"), ce.library_paths()[0], sysconfig.get_paths()["include"],
      sysconfig.get_config_var("EXT_SUFFIX"), int(torch._C._GLIBCXX_USE_CXX11_ABI))
PYEOF
)"
TORCH_INC=${TORCH_INC//;/ }
SRC="$REF/my_ext/_C"
FILES="src/pybind11.cpp src/nerf/gaussian_preprocess.cu src/nerf/gaussian_preprocess_colmap.cu \
src/nerf/gaussian_rasterizer_forward.cu src/nerf/gaussian_rasterizer_backwrad.cu \
src/nerf/gaussian_rasterizer_imp.cu src/nerf/gaussian_render.cu"

build_variant() {  # $1 = module name, $2 = extra nvcc flags
  local NAME="$1" EXTRA="$2"
  local TARGET="$OUT/$NAME$EXT_SUFFIX"
  if [ -f "$TARGET" ] && [ -z "${FORCE:-}" ]; then
    echo "[build_ref] $TARGET exists (FORCE=1 to rebuild)"; return 0
  fi
  local OBJ="$OUT/obj_$NAME"
  mkdir -p "$OBJ"
  local COMMON="-O3 -std=c++17 -I$SRC/include -I$SRC/third_party/glm $TORCH_INC -I$PY_INC \
 -DTORCH_EXTENSION_NAME=$NAME -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$CXX11_ABI \
 -rdc=true -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC --expt-relaxed-constexpr -w $EXTRA"
  local OBJS=""
  for f in $FILES; do
    local o="$OBJ/$(basename "${f%.*}").o"
    OBJS="$OBJS $o"
    echo "[build_ref] nvcc $EXTRA $f"
    nvcc -x cu $COMMON -c "$SRC/$f" -o "$o" &
  done
  wait
  nvcc -shared -rdc=true -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC $OBJS -L"$TORCH_LIB" -lc10 -lc10_cuda \
    -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -Xlinker -rpath -Xlinker "$TORCH_LIB" -o "$TARGET"
  rm -rf "$OBJ"
  echo "[build_ref] built $TARGET"
}

mkdir -p "$OUT"
build_variant _ref_raster ""
build_variant _ref_raster_nofma "-fmad=false"
