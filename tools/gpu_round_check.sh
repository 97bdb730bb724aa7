# One GPU-box pass: GPU tests, smoke, bench lines (ours + the reference arm), ncu launch lists and full captures of the
# dominant kernels.  usage (from the repo root, under gpurun): bash tools/gpu_round_check.sh <tag>
set -x
TAG=${1:-r02}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_reference_arm_$TAG.json 2> gpurun_out/bench_c2_reference_arm_$TAG.err
python bench.py --workload c1 --steps 10 --warmup 3 > gpurun_out/bench_c1_$TAG.json 2> gpurun_out/bench_c1_$TAG.err
if [ -z "$NO_NCU" ]; then
NCU="ncu --profile-from-start off --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-c1"
$B > gpurun_out/plain_c2_$TAG.log 2>&1 && $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_c2_$TAG.csv $B > gpurun_out/ncu_c2_$TAG.log 2>&1
$B --workload c1 > gpurun_out/plain_c1_$TAG.log 2>&1 && $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_c1_$TAG.csv $B --workload c1 > gpurun_out/ncu_c1_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:scan_mma_kernel -s 1 -c 1 -o gpurun_out/prof_scan_c2_$TAG -f $B > gpurun_out/ncu_full_c2_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:rerank_cta_kernel -s 1 -c 1 -o gpurun_out/prof_rerank_c2_$TAG -f $B > gpurun_out/ncu_rr_c2_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:approx_gemm_tc5 -s 1 -c 1 -o gpurun_out/prof_gemm_c1_$TAG -f $B --workload c1 > gpurun_out/ncu_gemm_c1_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:prefilter_select -s 1 -c 1 -o gpurun_out/prof_select_c1_$TAG -f $B --workload c1 > gpurun_out/ncu_sel_c1_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:scan_mma_kernel -s 1 -c 1 -o gpurun_out/prof_scan_c1_$TAG -f $B --workload c1 > gpurun_out/ncu_full_c1_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:quantize -s 1 -c 1 -o gpurun_out/prof_quant_c1_$TAG -f $B --workload c1 > gpurun_out/ncu_q_c1_$TAG.log 2>&1
fi
if [ -n "$SWEEPS" ]; then
python bench.py --steps 20 --warmup 5 --no-c1 --no-cpu --sweep-probe 16,32,64,128,256 --sweep-batch 1,8,64,512,1000 > gpurun_out/bench_c2_sweeps_$TAG.json 2> gpurun_out/bench_c2_sweeps_$TAG.err
python bench.py --workload c3 --steps 10 --warmup 3 --no-c1 --truth-queries 2000 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-c1 --truth-queries 1000 --cpu-seconds 8 > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
fi
tail -c 300 gpurun_out/bench_c2_$TAG.err
