set -x
mkdir -p gpurun_out
timeout 600 python scripts/bench_rollout.py --json gpurun_out/rollout.json > gpurun_out/rollout.log 2>&1
CA_STORE_MODE=vec4 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_vec4.json 2> gpurun_out/bench_x.err
CA_ONESHOT_MINBLOCKS=8 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_mb8.json 2>> gpurun_out/bench_x.err
CA_ONESHOT_MINBLOCKS=6 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_mb6.json 2>> gpurun_out/bench_x.err
CA_DISABLE_L2_PREFETCH=1 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_nopf.json 2>> gpurun_out/bench_x.err
CA_DISABLE_PDL=1 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_nopdl.json 2>> gpurun_out/bench_x.err
cat gpurun_out/rollout.log; for f in vec4 mb8 mb6 nopf nopdl; do cut -c1-230 gpurun_out/bench_$f.json; done
