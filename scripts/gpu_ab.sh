mkdir -p gpurun_out; rm -f gpurun_out/tune_ab.log
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q 2>&1 | tail -1 >> gpurun_out/tune_ab.log
for d in f64 f32; do for n in 256 512; do
  echo "-- old (61da503)" >> gpurun_out/tune_ab.log
  (cd tuning_libs/wt_old && timeout 300 python scripts/tune.py $n $d "kernel_variant=0") >> gpurun_out/tune_ab.log 2>&1
  echo "-- head" >> gpurun_out/tune_ab.log
  timeout 300 python scripts/tune.py $n $d "kernel_variant=0" >> gpurun_out/tune_ab.log 2>&1
  TUNE_RUN=20 timeout 300 python scripts/tune.py $n $d "kernel_variant=0,fused_step=0" "kernel_variant=4,fused_step=1" >> gpurun_out/tune_ab.log 2>&1
done; done
cat gpurun_out/tune_ab.log
