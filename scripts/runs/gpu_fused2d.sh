mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gradients.py tests/test_gpu_scaled_configs.py tests/test_gpu_fullsize.py tests/test_gpu_variants.py -q 2>&1 | tail -3
C5_STEPS=20000 timeout 900 python scripts/bench_configs.py c5 > gpurun_out/c5_20000_fused.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/c5_20000_fused.log"):
    if l.startswith("{"):
        r = json.loads(l); print(r["dtype"], "incl. tangents %.1f Gcell/s in %.2f s; primal alone %.1f" % (r["gcell_per_s_incl_tangents"], r["seconds"], r["primal_only_gcell_per_s"]))
PY
