# builds libpcbgpu variants with different launch-bound / CTA-size macros into variants/
# usage: bash scripts/build_variants.sh "name:flags" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
SRC="rl-environment-for-component-placement_b200/csrc"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -shared -Xcompiler -fPIC \
     $flags -o variants/libpcbgpu_$name.so $SRC/pcb_api.cu $SRC/pcb_gen.cpp 2>&1 | grep -E "error" || true
  echo "built $name ($flags)"
done
