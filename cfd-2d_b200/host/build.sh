#!/bin/bash
# Build the drop-in driver: reference host code (compiled from the sources where they lie under
# $REF, nothing copied) + the FVM_TVD_CUDA glue + libcfd2d_b200.so.  Needs the reference tree, so it
# only runs in the build container; the binary (git-ignored) travels to the GPU box.
set -e
cd "$(dirname "$0")"
REF=${REF:-/root/reference}
SRC=$REF/src
[ -f "$SRC/methods/fvm_tvd.cpp" ] || { echo "reference tree not found at $REF"; exit 1; }
mkdir -p _build/obj
CXX=${CXX:-g++}
FL="-std=c++11 -O2 -fPIC -fpermissive -w -ffp-contract=off -Impi_shim -I$SRC -I$SRC/methods -I$SRC/mesh -I$SRC/tinyxml"
for f in global bnd_cond mesh/grid mesh/MeshReader mesh/MeshReaderBerkleyTriangle mesh/MeshReaderSalomeUnv \
         methods/fvm_tvd tinyxml/tinystr tinyxml/tinyxml tinyxml/tinyxmlerror tinyxml/tinyxmlparser; do
  o=_build/obj/$(echo $f | tr / _).o
  [ -f $o ] || $CXX $FL -c $SRC/$f.cpp -o $o &
done
wait
$CXX $FL -c fvm_tvd_cuda.cpp -o _build/obj/glue.o
$CXX $FL -c main_cuda.cpp -o _build/obj/main.o
$CXX -o _build/cfd2d_cuda _build/obj/*.o -L../csrc -lcfd2d_b200 -Wl,-rpath,'$ORIGIN/../../csrc' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -lm
echo "built $(pwd)/_build/cfd2d_cuda"
