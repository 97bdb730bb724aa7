set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_c2_r01g.json 2> gpurun_out/bench_c2_r01g.err
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r01g.json 2> gpurun_out/bench_c1_r01g.err
