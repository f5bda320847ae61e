#!/bin/bash
# tests + step profile + bench + (optional) ncu full capture of selected kernels.  usage: gpu_round.sh <tag> [ncu-regex] [ncu-count]
TAG=${1:-x}; RE=${2:-}; CNT=${3:-12}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 200 python scripts/dev/step_prof.py > gpurun_out/${TAG}_stepprof.log 2>&1; echo "stepprof rc=$?"; head -40 gpurun_out/${TAG}_stepprof.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json")); print("bench ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"]); print(d["roofline"])
PY
if [ -n "$RE" ]; then
  CMD="python scripts/dev/step_prof.py 4096 rk4"
  ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 60 -c $CNT -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
  echo "ncu rc=$?"
fi
