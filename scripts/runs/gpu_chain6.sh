mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; tail -2 gpurun_out/smoke_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2_launches_f64_256.csv python bench.py --steps 2 --warmup 1 --chunk 100 --no-cpu --no-extra --no-e2e > gpurun_out/bench_ncu.log 2>&1; tail -c 300 gpurun_out/bench_ncu.log
python - <<PY
import csv, collections
for f in ("gpurun_out/r2_launches_smoke.csv", "gpurun_out/r2_launches_f64_256.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    t = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[4].split("(")[0][:70]
        t[name][0] += 1; t[name][1] += float(r[-1].replace(",", ""))
    print(f)
    for k, (n, ns) in sorted(t.items(), key=lambda kv: -kv[1][1])[:12]:
        print("   %-72s x%-4d %10.1f us" % (k, n, ns / 1e3))
PY
echo "== batched tangent sweep: x-chunk length"
for d in f32 f64; do TUNE_B=16 timeout 600 python scripts/tune2d.py 2048 $d "" xchunk=8 xchunk=16 xchunk=64 2>&1 | grep -v Warn | tail -4; done | tee gpurun_out/chain6_tune2d_chunk.log
