# round-2 closing measurements on one B200 (gpurun): GPU test suite, the driver-shaped bench lines, per-kernel shares,
# the ncu launch list of one step, and the DRAM-traffic capture of the stand-alone point operators
O=gpurun_out/r2z
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $O/pytest.log; tail -2 $O/pytest.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-160 $O/bench_default.json
timeout 600 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; cut -c1-200 $O/bench_reference_arm.json
timeout 600 python bench.py --masker gmm --no-extras > $O/bench_gmm_masker.json 2> $O/bench_gmm.err; cut -c1-160 $O/bench_gmm_masker.json
timeout 600 python bench.py --config 2 > $O/bench_config2_seg16384.json 2> $O/bench_config2.err; cut -c1-160 $O/bench_config2_seg16384.json
timeout 600 python bench.py --config 4 > $O/bench_config4_stress65536.json 2> $O/bench_config4.err; cut -c1-160 $O/bench_config4_stress65536.json
timeout 300 python bench.py --batch 64 --streams 3 --no-extras --no-cpu-baseline --shares-out $O/kernel_shares_b64.json > $O/bench_b64_s3.json 2>/dev/null; cut -c1-160 $O/bench_b64_s3.json
timeout 300 python bench.py --no-extras --no-cpu-baseline --shares-out $O/kernel_shares_b128.json > /dev/null 2>&1
timeout 200 python bench.py --batch 64 --streams 1 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > $O/bench_plain.json 2> $O/bench_plain.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/launches_bench_b64.csv python bench.py --batch 64 --streams 1 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 300 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file $O/point_ops_traffic.csv python scripts/prof_point_ops.py --order-out $O/point_ops_order.json > $O/prof_point_ops.log 2>&1
tail -2 $O/prof_point_ops.log
ls -la $O
