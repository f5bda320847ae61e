#!/bin/bash
# Round profile capture: bench line (with CPU baseline), ncu launch list, ncu --set full of the dominant kernels.
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
P="python scripts/dev/step_prof.py 4096 rk4"
$P > gpurun_out/plain2.log 2>&1 || exit 1
for spec in "k_chain_fwd:10:chainfwd" "k_chain_bwd:10:chainbwd" "k_gemm_k128_rows:10:y1rows" "k_gemm_tc:10:z0"; do
  IFS=: read K S NAME <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o gpurun_out/${TAG}_full_$NAME $P > gpurun_out/ncu_$NAME.log 2>&1
  echo "ncu $NAME rc=$?"
done
ls -la gpurun_out/*.ncu-rep
cut -c1-400 gpurun_out/${TAG}_bench.json
