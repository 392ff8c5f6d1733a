# builds libpmb.so variants with PMB_PAR_DBG=$1 into /tmp and times variant 8/9
set -e
cd /root/repo/pymoto_b200
for d in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --fmad=true -cudart shared -DPMB_PAR_DBG=$d csrc/pmb_api.cu csrc/pmb_assembly.cu csrc/pmb_spmv.cu csrc/pmb_multigrid.cu csrc/pmb_vector.cu csrc/pmb_filter.cu csrc/pmb_elem.cu csrc/pmb_optim.cu csrc/pmb_solver.cu csrc/pmb_probe.cu csrc/pmb_comm.cu -ldl -o ../gpurun_dbg/libpmb_dbg$d.so &
done
wait
