set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 1200 --warmup 48 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_ref.json 2>> gpurun_out/bench_8gpu.err
cut -c1-300 gpurun_out/bench_8gpu.json; cut -c1-200 gpurun_out/bench_8gpu_ref.json; tail -3 gpurun_out/bench_8gpu.err
