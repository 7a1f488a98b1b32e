mkdir -p gpurun_out
CA_EARLY_PREFETCH=1 timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_ep1.json 2>> gpurun_out/bench_x.err
timeout 200 python bench.py --no-cpu-baseline --steps 1200 > gpurun_out/bench_ep0.json 2>> gpurun_out/bench_x.err
CA_EARLY_PREFETCH=1 timeout 200 python bench.py --no-cpu-baseline --steps 1200 --workload phase2 > gpurun_out/bench_ep1_p2.json 2>> gpurun_out/bench_x.err
for f in ep1 ep0 ep1_p2; do echo "$f $(grep -o 'ms_per_step[^,]*' gpurun_out/bench_$f.json | head -1)"; done
