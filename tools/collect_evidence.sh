#!/bin/bash
# One GPU-box pass that produces everything tools/refresh_profiles.py turns into profiles/: run it through
#   gpurun --timeout 1500 -- 'bash tools/collect_evidence.sh r1'
# Bench numbers come from plain runs; the ncu passes afterwards only provide launch shares and counters.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
python bench.py --workload small --no-e2e > $OUT/bench_small.json 2> $OUT/bench_small.err
python bench.py --workload ippo --steps 3 > $OUT/bench_ippo.json 2> $OUT/bench_ippo.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:env_place|env_alloc|env_feature|env_reward|env_step|env_reset|gae_kernel|moments|standardize' -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu_list.log 2>&1
# one launch of each kernel of a steady-state step (t = 35 of the third warm-up segment's episode)
ncu --set full --import-source on --clock-control none -k 'regex:env_place|env_alloc|env_feature' --launch-skip 405 --launch-count 3 \
    -f -o $OUT/${TAG}_k1_split python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1
cat $OUT/${TAG}_pytest_gpu.log
head -c 600 $OUT/bench_default.json
