mkdir -p gpurun_out; rm -f gpurun_out/v5d_*.log
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q -k "tensor_map" > gpurun_out/v5d_pytest.log 2>&1
tail -3 gpurun_out/v5d_pytest.log
for d in f64 f32; do for n in 256 512; do
  timeout 600 python scripts/tune.py $n $d "kernel_variant=0" "kernel_variant=6,tma_rows=4,tma_stages_H=3,tma_stages_D=4" "tma_stages_D=3" \
     "tma_stages_H=3,tma_stages_D=3,xchunk=8" "tma_stages_D=4" >> gpurun_out/v5d_tune.log 2>&1
done; done
cat gpurun_out/v5d_tune.log
for k in H D; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step_${k}_v5 -s 5 -c 1 -o gpurun_out/prof_v5d_${k}_f64_256 \
   python scripts/tune.py 256 f64 "kernel_variant=6,tma_rows=4,tma_stages_H=3,tma_stages_D=4" > gpurun_out/v5d_ncu.log 2>&1
done
tail -3 gpurun_out/v5d_ncu.log
