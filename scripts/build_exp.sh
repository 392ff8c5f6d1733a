#!/bin/bash
# Diagnostic builds of libpmb.so with layout experiments compiled in (PMB_PAR_EXP bit mask, pmb_elem_par.cuh): every source but
# pmb_elem.cu is compiled once into /tmp/pmbobj, pmb_elem.cu once per experiment; libraries land in scripts/_exp/ (git-ignored,
# shipped to the GPU box) and are selected with PMB_LIB_PATH.   usage: scripts/build_exp.sh 0 1 2 ...
set -e
cd "$(dirname "$0")/.."
OBJ=/tmp/pmbobj; mkdir -p $OBJ scripts/_exp
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=true -Xcompiler -fPIC -cudart shared"
for f in pymoto_b200/csrc/*.cu; do
  b=$(basename $f .cu); [ $b = pmb_elem ] && continue
  if [ ! -f $OBJ/$b.o ] || [ $f -nt $OBJ/$b.o ]; then nvcc $FLAGS -c $f -o $OBJ/$b.o & fi
done
for e in "$@"; do nvcc $FLAGS -DPMB_PAR_EXP=$e -c pymoto_b200/csrc/pmb_elem.cu -o $OBJ/pmb_elem_exp$e.o & done
wait
for e in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart shared $(ls $OBJ/*.o | grep -v pmb_elem_exp) $OBJ/pmb_elem_exp$e.o -ldl -o scripts/_exp/libpmb_exp$e.so
done
ls -la scripts/_exp/
