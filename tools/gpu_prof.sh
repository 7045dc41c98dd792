#!/bin/bash
# One full ncu capture of the saturating search launch; the report comes back in gpurun_out/ (read here with
# tools/ncu_summary.py / tools/ncu_phases.py).  Usage: tools/gpu_prof.sh <tag> [n] [nq]
TAG=${1:-prof}; N=${2:-1000000}; NQ=${3:-32768}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "prof/" \
    -k regex:search_kernel -c 1 -f -o gpurun_out/${TAG} python tools/prof_search.py $N $NQ 3 > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}.ncu-rep
