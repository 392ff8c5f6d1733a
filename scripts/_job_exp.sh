# time the PMB_PAR_EXP builds of the parity-block kernel (scripts/build_exp.sh) against each other
mkdir -p gpurun_out
for e in "$@"; do
  PMB_LIB_PATH=$PWD/scripts/_exp/libpmb_exp$e.so timeout 200 python scripts/time_elem.py --variants ${VARIANTS:-0,8,9} --cases 3:256x128x128,1:256x256x256 --reps 30 --out gpurun_out/time_elem_exp$e.json > /dev/null 2> gpurun_out/time_elem_exp$e.err || tail -5 gpurun_out/time_elem_exp$e.err
  python - <<P
import json
d=json.load(open("gpurun_out/time_elem_exp$e.json"))
print("exp $e", {c: {v: (round(o["ms"],4), "%.1e" % o["maxdiff"]) for v,o in r.items()} for c,r in d.items()})
P
done
