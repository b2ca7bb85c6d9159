#!/bin/bash
# usage: scripts/ptxas_info.sh file.cu [filter]  -- registers / spills per kernel (c++filt names)
cd "$(dirname "$0")/../prlib_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -c "$1" -o /tmp/ptxas_info.o 2>&1 \
 | awk '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)} /spill/ {sp=$0} /Used/ {print name " | " $0 " | " sp}' \
 | c++filt | sed -e 's/ptxas info    ://g' -e 's/(anonymous namespace):://g' | cut -c1-330 | grep -E "${2:-.}"
