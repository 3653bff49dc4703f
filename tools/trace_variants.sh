#!/bin/bash
# run tools/tc_trace.py on several library variants; prints the member length (cycles) of each
for n in "$@"; do
  NPHM_B200_LIB=$PWD/nphm_b200/libnphm_b200_$n.so timeout 200 python tools/tc_trace.py > gpurun_out/tc_trace_$n.txt 2>&1
  echo "$n: $(grep -m3 'member length' gpurun_out/tc_trace_$n.txt | sed 's/.*member length \([0-9]*\)).*/\1/' | tr '\n' ' ')"
done
