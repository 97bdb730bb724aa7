set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_r01m.json 2> gpurun_out/bench_c1_r01m.err
python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c3_r01m.json 2> gpurun_out/bench_c3_r01m.err
