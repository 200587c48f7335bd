#!/bin/bash
rm -f gpurun_out/count_ab.log
export KMU_COUNT_TIMING=1
for mode in twophase direct; do
  if [ $mode = direct ]; then export KMU_COUNT_DIRECT=1; else unset KMU_COUNT_DIRECT; fi
  echo "== $mode reads=8000000" >> gpurun_out/count_ab.log
  timeout 300 python scripts/bench_count.py --reads 8000000 --rounds 3 2>&1 | tail -9 >> gpurun_out/count_ab.log
done
cat gpurun_out/count_ab.log
