#!/bin/bash
# vkhrt_headless (C++ host mirror, vkhrt_render_multi from one process) at N = 1..8 on bench.py's weak-scaling frames, hit records only
L=vkhrt_b200/_lib
declare -A SZ=( [1]=1920x1080 [2]=2720x1530 [4]=3840x2160 [8]=5432x3056 )
for n in ${*:-1 2 4 8}; do
  $L/vkhrt_headless --model synthetic:curly:100000:32 --technique phantom --size ${SZ[$n]} --frames 12 --gpus $n --no-image 2>&1 | awk -v n=$n -v sz=${SZ[$n]} '/^frame/ {c++; if (c>4) {s+=$3; k++}} END {split(sz,a,"x"); printf "headless --gpus %d %s: %.3f ms/frame wall (mean of frames 5-12) = %.1f Mrays/s e2e\n", n, sz, s/k, a[1]*a[2]/(s/k)/1e3}'
done
