#!/bin/bash
# builds tests/native/umma_probe (stand-alone tcgen05 descriptor probe)
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe umma_probe.cu
