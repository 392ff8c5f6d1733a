for d in a b d e f; do
PMB_LIB_PATH=$PWD/gpurun_dbg/libpmb_$d.so timeout 300 python scripts/time_sym.py --out gpurun_out/time_sym_$d.json > gpurun_out/time_sym.log 2>&1; echo "cfg=$d" $(grep -E "\"ms\"|maxdiff" gpurun_out/time_sym_$d.json | tr -d ' \n')
done
