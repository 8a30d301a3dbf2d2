# D half-step of the tensor-map kernels with a register cap for 4 CTAs per SM (-DCEV_V5_D_MINB=4: 128 registers in fp64, no
# spills) x ring depth (3 stages = 44 KB per CTA lets 4 CTAs fit; 4 stages = 58 KB allows 3), against the default build
mkdir -p gpurun_out; out=gpurun_out/r2_tune_v5_D_register_cap.log; : > $out
A=$PWD/ceviche_b200/libceviche_b200.so; B=$PWD/ceviche_b200/libceviche_b200_dcap4.so
for n in 256 512; do
  echo "-- default build (D: 158 registers fp64)" >> $out
  CEV_LIB_PATH=$A timeout 100 python scripts/tune.py $n f64 "tma_stages_D=4" "tma_stages_D=3" >> $out 2>&1
  echo "-- -DCEV_V5_D_MINB=4 (D: 128 registers fp64)" >> $out
  CEV_LIB_PATH=$B timeout 100 python scripts/tune.py $n f64 "tma_stages_D=4" "tma_stages_D=3" >> $out 2>&1
done
cat $out
CEV_LIB_PATH=$B timeout 200 python -m pytest tests/test_gpu_variants.py -q -k "tensor_map" 2>&1 | tail -2
