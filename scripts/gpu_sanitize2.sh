# compute-sanitizer over the round-2 kernels (tensor-map adjoint, batched tangents, recorder, fold, graph replay)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gradients.py -x -q -k "tensor_map or graphs or forward_mode or batched" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_gradients.py -x -q -k "tensor_map_adjoint_identity" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scaled_configs.py -x -q -k "config5" >> gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck c5 rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
