#!/bin/bash
# The single-GPU C3 counting step on every GPU of the box in turn, and then on all of them at once as independent
# single-GPU jobs: separates "this GPU is slower at random table updates" (differences already in the first pass) from
# "the GPUs slow each other down when all are busy" (differences only in the second) -- the open question behind the
# per-rank spread of the 8-GPU insertion.  Needs a box with N GPUs: scripts/per_gpu_c3.sh 8
set -u
n=${1:-8}
out=${2:-gpurun_out/per_gpu_c3}
mkdir -p "$out"
pick='import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for c in d["configs"]:
    print(round(c["value"],1), "Gbases/s", round(c["ms_per_step"],1), "ms", c.get("phases_ms_per_step"))'
echo "== one at a time"
for g in $(seq 0 $((n - 1))); do
    CUDA_VISIBLE_DEVICES=$g python bench.py --configs c3 --steps 2 --warmup 1 2> "$out/solo_$g.err" | tee "$out/solo_$g.json" | python -c "$pick" | sed "s/^/gpu $g: /"
done
echo "== all at once (independent single-GPU jobs)"
for g in $(seq 0 $((n - 1))); do
    CUDA_VISIBLE_DEVICES=$g python bench.py --configs c3 --steps 2 --warmup 1 > "$out/together_$g.json" 2> "$out/together_$g.err" &
done
wait
for g in $(seq 0 $((n - 1))); do
    python -c "$pick" < "$out/together_$g.json" | sed "s/^/gpu $g: /"
done
