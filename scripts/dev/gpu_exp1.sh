#!/bin/bash
# developer probe: chain-kernel phase traces under several timing-experiment builds (results of the experiment builds are wrong by design)
mkdir -p gpurun_out
for v in "" "-DCHAIN_FAKE_RESIDENT" "-DCHAIN_NO_RING" "-DCHAIN_SKIP_BF16"; do
  echo "=== variant [$v] ===" 
  bash scripts/dev/gpu_trace.sh "$v" 30
done > gpurun_out/r2e_exp1.txt 2>&1
echo "=== grid 148 ===" >> gpurun_out/r2e_exp1.txt
cd swarm_ode_b200/csrc && touch chain_fwd.cu && make EXTRA="-DCHAIN_TRACE" -j8 > /tmp/mk.log 2>&1; cd ../..
CHAIN_GRID=148 timeout 300 python scripts/dev/chain_trace.py 2>&1 | head -14 >> gpurun_out/r2e_exp1.txt
CHAIN_GRID=148 timeout 120 python scripts/dev/step_prof.py 4096 rk4 2>&1 | grep -E "chain|^step" | head -4 >> gpurun_out/r2e_exp1.txt
