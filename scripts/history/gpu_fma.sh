mkdir -p gpurun_out; rm -f gpurun_out/tune_fma.log
for L in nofma fma; do echo "== lib $L" >> gpurun_out/tune_fma.log
 CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python -m pytest tests/test_gpu_fields.py tests/test_gpu_variants.py -x -q 2>&1 | tail -1 >> gpurun_out/tune_fma.log
 for d in f64 f32; do for n in 256 512; do CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 120 python scripts/tune.py $n $d "kernel_variant=0" >> gpurun_out/tune_fma.log 2>&1; done; done
 CEV_LIB_PATH=$PWD/tuning_libs/lib_$L.so timeout 300 python bench.py --no-cpu --no-e2e --steps 6 > gpurun_out/bench_$L.json 2>/dev/null
 python -c "
import json; d=json.load(open('gpurun_out/bench_$L.json')); print('sustained', round(d['value'],2), d['clocks'])" >> gpurun_out/tune_fma.log
done
cat gpurun_out/tune_fma.log
