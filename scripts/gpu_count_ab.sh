#!/bin/bash
python -m pytest tests/test_count_twophase_gpu.py tests/test_count_gpu.py -x -q -m gpu > gpurun_out/count_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/count_tests.log
tail -3 gpurun_out/count_tests.log
rm -f gpurun_out/count_ab.log
export KMU_COUNT_TIMING=1
for l1 in 512 128 64 32; do
  export KMU_COUNT_LEVEL1_BUCKETS=$l1
  echo "== level1=$l1 reads=26666667" >> gpurun_out/count_ab.log
  timeout 300 python scripts/bench_count.py --reads 26666667 --rounds 2 2>&1 | grep -E "kmu count\] chunk|gbases" | tail -5 >> gpurun_out/count_ab.log
done
cat gpurun_out/count_ab.log
