#!/bin/bash
# developer probe: A/B of two builds of one source file inside ONE box (run-to-run variance between boxes is ~5 %)
# usage: gpu_ab.sh <file.cu> "<defines A>" "<defines B>" "<python command>"
F=$1; A=$2; B=$3; CMD=$4
for round in 1 2; do
  for v in A B; do
    D=$A; [ $v = B ] && D=$B
    (cd swarm_ode_b200/csrc && touch $F && make EXTRA="$D" -j8 > /tmp/mk.log 2>&1 || tail -5 /tmp/mk.log)
    echo "=== round $round variant $v [$D] $(nvidia-smi --query-gpu=clocks.sm,temperature.gpu --format=csv,noheader)"
    eval "$CMD"
  done
done
