#!/bin/bash
# usage: ptxas_report.sh file.cu [pattern]  -- registers / spills / smem per kernel (no GPU needed)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xptxas -v -c "$1" -o /dev/null 2>&1 \
 | grep -E "error|Function properties|Used|spill" | sed -e 's/ptxas info    : //' | paste - - - 2>/dev/null | grep -E "${2:-.}" | cut -c1-260 | head -${3:-60}
