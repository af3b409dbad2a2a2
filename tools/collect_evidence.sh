#!/bin/bash
# One GPU-box pass that produces everything tools/refresh_profiles.py turns into profiles/: run it through
#   gpurun --timeout 1800 -- 'bash tools/collect_evidence.sh r2'
# Bench numbers come from plain runs; the ncu passes afterwards only provide launch shares and counters.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $OUT/${TAG}_pytest_gpu.log
python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
python bench.py --layout wide --no-e2e --no-cpu > $OUT/bench_wide.json 2> $OUT/bench_wide.err
python bench.py --workload small --no-e2e > $OUT/bench_small.json 2> $OUT/bench_small.err
python bench.py --workload ippo --steps 3 > $OUT/bench_ippo.json 2> $OUT/bench_ippo.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:compact_|env_place|env_alloc|env_feature|env_reward|env_step|env_reset|lines_from|gae_kernel|moments|standardize' -c 500 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-spot-check > $OUT/${TAG}_ncu_list.log 2>&1
# one launch of each kernel of a steady-state step (t = 35 of the recording episode)
STEPS=40 PLAIN=1 ncu --set full --import-source on --clock-control none -k 'regex:compact_place|compact_alloc|compact_feature' --launch-skip 105 --launch-count 3 \
    -f -o $OUT/${TAG}_k1_compact python tools/step_timings.py > $OUT/${TAG}_ncu_full.log 2>&1
STEPS=40 python tools/step_timings.py > $OUT/${TAG}_step_timings.log 2>&1
python tools/pipeline_timings.py > $OUT/${TAG}_pipeline_timings.log 2>&1
ENVS="4096 16384 32768 262144" TEAMS="1 8" bash tools/small_team_sweep.sh > $OUT/${TAG}_small_team_sweep.log 2>&1
cat $OUT/${TAG}_pytest_gpu.log
head -c 600 $OUT/bench_default.json
