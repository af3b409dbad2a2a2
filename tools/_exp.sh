python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
python bench.py --workload ippo --steps 2 --warmup 2 2>/dev/null | cut -c1-330
python bench.py --workload mappo --steps 2 --warmup 2 2>/dev/null | cut -c1-330
