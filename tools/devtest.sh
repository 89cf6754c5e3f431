#!/bin/bash
# Runs every devtest case in its own process (a trapped kernel poisons the CUDA context).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=$(./build/devtest list)
: > gpurun_out/devtest.log
for i in $(seq 0 $((n-1))); do
  timeout 60 ./build/devtest $i >> gpurun_out/devtest.log 2>&1
  echo "[case $i] exit=$?" >> gpurun_out/devtest.log
done
grep -c PASS gpurun_out/devtest.log
grep -E "^\[case [0-9]+\] (PASS|FAIL)|exit=" gpurun_out/devtest.log | paste - - | head -60
