#!/bin/bash
# usage: sass_stats.sh <file.cu (in csrc/)>: compile one kernel file for sm_100a with -Xptxas -v and print the SASS opcode histogram
P=/root/repo/structure-from-motion-3d-reconstruction_b200
F=$1
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v -c $P/csrc/$F.cu -o $P/build/$F.o 2>&1 | grep -E "Compiling|Used|spill|error|warning"
cuobjdump -sass $P/build/$F.o | grep -oE "^\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${2:-14}
