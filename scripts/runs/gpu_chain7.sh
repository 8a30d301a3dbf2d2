mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gradients.py -q -k "forward_mode or reference_forward" > gpurun_out/chain7_pytest.log 2>&1; tail -5 gpurun_out/chain7_pytest.log | cut -c1-300
for o in "" "tma_stages_adjH=3" "tma_stages_adjED=4" "tma_stages_adjH=3,tma_stages_adjED=4"; do
  echo "-- adjoint ring depths: '$o'"
  C4_OPTS="$o" C4_STEPS=400 timeout 600 python scripts/bench_configs.py c4 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print('  ', r['dtype'], r['config'][60:100], 'bwd %.1f' % r['backward_gcell_per_s'])"
done | tee gpurun_out/chain7_adj_stages.log
