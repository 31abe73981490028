#!/bin/bash
mkdir -p gpurun_out
g++ -O3 -march=native -fopenmp scripts/micro/host_expand2.cpp -o /tmp/host_expand2 && /tmp/host_expand2 65536 | tee gpurun_out/host_expand2.log
python scripts/micro/pcie_bw.py 2>&1 | tail -6 | tee -a gpurun_out/host_expand2.log
