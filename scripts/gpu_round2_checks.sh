mkdir -p gpurun_out/r2b
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2b/pytest.log
tail -c 1500 gpurun_out/r2b/pytest.log
timeout 300 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b/point_ops_traffic.csv python scripts/prof_point_ops.py --order-out gpurun_out/r2b/point_ops_order.json > gpurun_out/r2b/prof_point_ops.log 2>&1
tail -3 gpurun_out/r2b/prof_point_ops.log
# sanitizer: memcheck + racecheck over a reduced subset (tensor-core kernels, block kNN, cluster FPS, CSR, dataset kernels)
SUB="tests/test_gpu_tc.py tests/test_gpu_point_ops.py::test_reference_twins_golden tests/test_gpu_point_ops.py::test_knn_blocks_equals_brute_force tests/test_gpu_point_ops.py::test_segment_softmax_sum_fused tests/test_host.py::test_carla_subsampler_on_device_matches_reference_golden tests/test_gpu_model.py::test_tflow_end_to_end_vs_reference_golden tests/test_gpu_frontend.py::test_batched_frontend_equals_single"
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest $SUB "tests/test_gpu_point_ops.py::test_fps" -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2b/sanitizer_memcheck.log
tail -8 gpurun_out/r2b/sanitizer_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py::test_tflow_end_to_end_vs_reference_golden -m gpu -q -x -k "not 8192" 2>&1 | tail -30) > gpurun_out/r2b/sanitizer_racecheck.log
tail -8 gpurun_out/r2b/sanitizer_racecheck.log
