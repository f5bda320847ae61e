#!/bin/bash
# Round-1 baseline capture: GPU parity tests, bench line, ncu launch list and one full capture of the top kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_sgemm' -s 30 -c 6 -o gpurun_out/r1_prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/r1_pytest_gpu.log
cat gpurun_out/r1_bench.json | cut -c1-600
