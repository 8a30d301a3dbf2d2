# fused tangent step (tan2d_fused.cuh): L2 prefetch distance x register cap (CTAs per SM), config 5's grid, 2000 steps
mkdir -p gpurun_out; out=gpurun_out/r2_tune_fused_tangent_occupancy.log; : > $out
run() {  # label, lib, opts
  echo "-- $1 C5_OPTS='$3'" >> $out
  CEV_LIB_PATH=$2 C5_OPTS="$3" C5_STEPS=2000 timeout 120 python scripts/bench_configs.py c5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r = json.loads(l); print('   %s incl. tangents %.1f Gcell/s, %.3f s' % (r['dtype'], r['gcell_per_s_incl_tangents'], r['seconds']))
" >> $out
}
A=$PWD/ceviche_b200/libceviche_b200.so; B=$PWD/ceviche_b200/libceviche_b200_minb5.so
run "4 CTAs/SM" $A ""
run "4 CTAs/SM" $A "prefetch_planes=3"
run "4 CTAs/SM" $A "prefetch_planes=6"
run "5 CTAs/SM" $B ""
run "5 CTAs/SM" $B "prefetch_planes=3"
run "4 CTAs/SM" $A ""
cat $out
