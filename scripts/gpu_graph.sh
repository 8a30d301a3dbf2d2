mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py -x -q -k "graph" 2>&1 | tail -15
timeout 600 python scripts/bench_configs.py c1 2>&1 | tee gpurun_out/configs_graph.log
timeout 300 python scripts/tune2d.py 640 f32 "use_graph=0" "use_graph=1" "use_graph=0" "use_graph=1" 2>&1 | tee -a gpurun_out/configs_graph.log
timeout 300 python scripts/tune2d.py 320 f64 "use_graph=0" "use_graph=1" "use_graph=0" "use_graph=1" 2>&1 | tee -a gpurun_out/configs_graph.log
