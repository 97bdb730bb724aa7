# One GPU-box pass: GPU tests, smoke, bench lines, ncu launch lists and full captures of the two dominant kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round_check.sh <tag>
set -x
TAG=${1:-r01c}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err
RABITQ_RR_PREFETCH=0 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_c2_${TAG}_nopf.json 2>/dev/null
RABITQ_RR_PREFETCH=0 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c1_${TAG}_nopf.json 2>/dev/null
if [ -z "$NO_NCU" ]; then
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_c2_$TAG.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c1_$TAG.csv python bench.py --workload c1 --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_c1_$TAG.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:^scan_kernel -s 1 -c 1 -o gpurun_out/prof_scan_c2_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_c2_$TAG.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:^rerank_kernel -s 1 -c 1 -o gpurun_out/prof_rerank_c2_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_rr_c2_$TAG.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:^scan_kernel -s 1 -c 1 -o gpurun_out/prof_scan_c1_$TAG -f python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_c1_$TAG.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:^rotate_kernel -c 1 -o gpurun_out/prof_rotate_c2_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_rot_c2_$TAG.log 2>&1
fi
tail -c 300 gpurun_out/bench_c2_$TAG.err
