mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
for sp in 0 1; do
  echo "== CEV_SPECIALISE=$sp"
  CEV_SPECIALISE=$sp timeout 600 python scripts/bench_configs.py c1 c5 2>&1 | tee -a gpurun_out/configs_mask_$sp.log
done
