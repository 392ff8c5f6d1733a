# usage: _dbg_build.sh name "-Dflags" ...   builds gpurun_dbg/libpmb_<name>.so (diagnostic builds, not tracked)
cd /root/repo/pymoto_b200
mkdir -p ../gpurun_dbg
while [ $# -gt 1 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --fmad=true -cudart shared -diag-suppress 128 $2 csrc/pmb_api.cu csrc/pmb_assembly.cu csrc/pmb_spmv.cu csrc/pmb_symstore.cu csrc/pmb_multigrid.cu csrc/pmb_vector.cu csrc/pmb_filter.cu csrc/pmb_elem.cu csrc/pmb_optim.cu csrc/pmb_solver.cu csrc/pmb_probe.cu csrc/pmb_comm.cu -ldl -o ../gpurun_dbg/libpmb_$1.so &
  shift; shift
done
wait
