#!/bin/bash
# Developer probe: rebuild chain_fwd.cu with different -D variants on the GPU box and print the chain kernel time of each.
cd swarm_ode_b200/csrc
for V in "$@"; do
  touch chain_fwd.cu
  make EXTRA="$V" > /tmp/mk.log 2>&1 || { tail -5 /tmp/mk.log; continue; }
  (cd ../.. && timeout 120 python scripts/dev/step_prof.py 4096 rk4 2>&1 | grep -E "^step:|chain_fwd" | tr '\n' ' '; echo " <= [$V]")
done
touch chain_fwd.cu; make > /tmp/mk.log 2>&1
