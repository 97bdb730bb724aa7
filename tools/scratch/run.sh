set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2_g8_r01n.json 2> gpurun_out/bench_c2_g8_r01n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_c2_g4_r01n.json 2> gpurun_out/bench_c2_g4_r01n.err
