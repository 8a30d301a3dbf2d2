mkdir -p gpurun_out; rm -f gpurun_out/tune10.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
for d in f64 f32; do for n in 256 512; do timeout 300 python scripts/tune.py $n $d "split_launch=0" >> gpurun_out/tune10.log 2>&1; done; done
cat gpurun_out/tune10.log
