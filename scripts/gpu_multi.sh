set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 2500 gpurun_out/bench_1gpu.json; tail -3 gpurun_out/bench_1gpu.err
