#!/bin/bash
# rebuild with in-kernel phase timestamps and print the chain-kernel traces (developer probe; the box is ephemeral)
# usage: gpu_trace.sh "<extra nvcc defines>" [fwd-only]
cd swarm_ode_b200/csrc && touch chain_fwd.cu chain_bwd.cu gemm_k128.cu && make EXTRA="-DCHAIN_TRACE $1" -j8 > /tmp/mk.log 2>&1 || { tail -20 /tmp/mk.log; exit 1; }
cd ../.. && timeout 300 python scripts/dev/chain_trace.py 2>&1 | head -${2:-40}
timeout 120 python scripts/dev/step_prof.py 4096 rk4 2>&1 | grep -E "chain|^step" | head -4
