#!/bin/bash
# Build libcfd2d_b200.so for sm_100a (B200).  -fmad=false: the reference build has no FMA contraction
# (SURVEY.md Appendix A), parity is judged against it.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2,-ffp-contract=off ${CFD2D_NVCC_EXTRA}"
$NVCC $FLAGS -Xptxas -v -c fvm_api.cu -o fvm_api.o 2> ptxas_fvm_api.log || { cat ptxas_fvm_api.log; exit 1; }
$NVCC $FLAGS -c halo_nccl.cu -o halo_nccl.o
${CXX:-g++} -O2 -std=c++17 -fPIC -c unv_reader.cpp -o unv_reader.o
OUT=${CFD2D_OUT:-libcfd2d_b200.so}
$NVCC -shared -o $OUT fvm_api.o halo_nccl.o unv_reader.o -lcudart -ldl
python3 - <<'PY' > ptxas_summary.txt || true
import re
out, name = [], None
for line in open("ptxas_fvm_api.log"):
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m: name, spill = m.group(1), ""; continue
    if name and "Function properties for " + name in line: own = True; continue
    if name and "spill" in line and not spill: spill = line.strip(); continue
    if name and "Used" in line and "registers" in line:
        out.append("%s\t%s\t%s" % (name, spill, line.replace("ptxas info    : ", "").strip())); name = None
print("\n".join(sorted(out)))
PY
echo "built $(pwd)/$OUT"
