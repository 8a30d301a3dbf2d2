mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
import ceviche_b200
from oracle import cases
from oracle.fdtd_numpy import OracleFDTD, rel_l2, FIELD_KEYS
rng = np.random.default_rng(0)
shape, npml, steps = (12, 10, 72), (3, 2, 5), 12
eps = 1 + rng.random(shape)
src = [("z", rng.random(shape) * (rng.random(shape) < 0.05), cases.gaussian(steps, 4, 3)), ("x", cases.one_hot(shape, (6, 5, 30)), cases.gaussian(steps, 5, 2))]
prb = [("Ez", rng.random(shape)), ("Hy", cases.one_hot(shape, (11, 9, 71)))]
O = OracleFDTD(eps, 5e-8, list(npml)); os_, _ = O.run(steps, src, prb)
for dtype in (torch.float64, torch.float32):
    for variant in (1, 2, 3, 0):
        F = ceviche_b200.fdtd(eps, 5e-8, list(npml), dtype=dtype)
        F.set_option("kernel_variant", variant)
        s = F.run(steps, src, prb)
        err = max(rel_l2(F.fields[k].cpu().numpy(), O.fields()[k]) for k in FIELD_KEYS)
        print(dtype, variant, "max rel-L2", err, flush=True)
# gradient sweep
e = torch.as_tensor(eps).cuda().requires_grad_(True)
F = ceviche_b200.fdtd(e, 5e-8, list(npml))
(F.run(steps, src, prb) ** 2).sum().backward()
print("grad norm", float(e.grad.norm()))
F2 = ceviche_b200.fdtd(eps, 5e-8, list(npml))
s, ds = F2.jvp_run(steps, torch.as_tensor(rng.standard_normal((2,) + shape)), src, prb)
print("jvp", float(ds.abs().max()))
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -15 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -8 gpurun_out/sanitizer_racecheck.log
timeout 600 python -m pytest tests/test_gpu_fields.py -x -q -k edge 2>&1 | tail -3
