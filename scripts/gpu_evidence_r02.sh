# Round-2 evidence run on one B200 (gpurun --timeout 2400 -- 'bash scripts/gpu_evidence_r02.sh'); raw outputs land in
# gpurun_out/ and are summarised into profiles/ by scripts/summarise_r02.py (run in the build container).
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/r02_bench_ref.json 2>> gpurun_out/r02_bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_k20.json 2>> gpurun_out/r02_bench.err
# launch list of the bench command (cold-cache, serialised per-launch times: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 96 --warmup 12 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
# steady-state DRAM traffic: no cache flush between the rotated launches, 24 consecutive launches
for wl in phase1 phase2 ragged; do
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum -k regex:ca_step_kernel -s 36 -c 24 --csv --log-file gpurun_out/r02_traffic_$wl.csv python bench.py --workload $wl --steps 96 --warmup 12 --no-cpu-baseline --no-extras --streams 1 > gpurun_out/r02_ncu_traffic_$wl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ca_step_kernel -s 40 -c 1 -o gpurun_out/r02c_step_$wl -f python bench.py --workload $wl --steps 96 --warmup 12 --no-cpu-baseline --no-extras --streams 1 > gpurun_out/r02_ncu_full_$wl.log 2>&1
done
if [ -n "$CA_EVIDENCE_PREDICTOR" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:predict_kernel -s 2 -c 1 -o gpurun_out/r02c_predict -f python scripts/predict_probe.py one 9 163840 0 > gpurun_out/r02_ncu_predict.log 2>&1
fi
timeout 300 python scripts/step_sweep.py phase1: phase1:_STREAMS=3 phase2: phase2:_STREAMS=3 ragged: ragged:_STREAMS=3 > gpurun_out/r02_sweep_final.log 2>&1
tail -2 gpurun_out/r02_smoke.log; tail -3 gpurun_out/r02_pytest_gpu.log; cut -c1-300 gpurun_out/r02_bench.json
