#!/bin/bash
# round-2 GPU pass.  usage: gpu_r2.sh <tag> [phases: t=tests b=bench r=reference-arm n=ncu-launch-list f=ncu-full(regex in $3)]
TAG=${1:-r2x}; PH=${2:-tb}; RE=${3:-}; CNT=${4:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader; nproc; free -g | head -2
if [[ $PH == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
fi
if [[ $PH == *b* ]]; then
  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "| e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"])
print("e2e_fp32", d.get("e2e_fp32_upload")); print("dd", d.get("e2e_device_dataset"))
print("roofline", d["roofline"]); print("parity", d.get("parity")); print("dopri5", d.get("dopri5_strong")); print("small", d.get("small_batch"))
for k in d["kernels"]: print(k)
print("cpu", d.get("cpu_baseline"))
PY
fi
if [[ $PH == *r* ]]; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
fi
if [[ $PH == *n* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-device-dataset --no-dopri5 --no-parity --no-side > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
if [[ $PH == *f* && -n "$RE" ]]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 40 -c $CNT -o gpurun_out/${TAG}_full \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-device-dataset --no-dopri5 --no-parity --no-side > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
