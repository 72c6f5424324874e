mkdir -p gpurun_out/r2f
# warm run without ncu must exit 0 first
timeout 200 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2f/bench_plain.json 2> gpurun_out/r2f/bench_plain.err || exit 1
# launch list of one step (times are cold-cache and serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2f/ncu_launch.log 2>&1
# full-set rows for every dense_tc launch of the warm-up forward + the two cost volumes
timeout 900 ncu --set full --clock-control none -k regex:"dense_tc_kernel|cost_volume_tc64" --launch-count 88 --csv --page raw --log-file gpurun_out/r2f/tc_full_raw.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2f/ncu_full.log 2>&1
ls -la gpurun_out/r2f
