#!/bin/bash
# One full ncu capture of the saturating search launch; the report comes back in gpurun_out/ (read here with
# tools/ncu_summary.py / tools/ncu_phase_lines.py / tools/ncu_sass.py).  Usage: tools/gpu_prof.sh <tag> [config] [nq]
TAG=${1:-prof}; CFG=${2:-c2}; NQ=${3:-32768}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "prof/" \
    -k regex:search_kernel -c 1 -f -o gpurun_out/${TAG} python tools/prof_search.py $CFG $NQ 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out/${TAG}.ncu-rep
