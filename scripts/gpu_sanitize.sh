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
# ---- kernel families added late in round 1
for dtype in (torch.float64, torch.float32):
    F = ceviche_b200.fdtd(eps, 5e-8, list(npml), dtype=dtype)          # fused full-step kernel, every tile shape
    for fs in (0, 804, 1608, 3204, 3208):
        F.initialize_fields()
        F.set_option("kernel_variant", 4); F.set_option("fused_shape", fs); F.set_option("xchunk", 5)
        F.run(steps + 1, [(c, p, np.append(w, 0.0)) for c, p, w in src], prb)
    print(dtype, "fused ok", flush=True)
    # 2-D grids: relabelled (x, z, y), compile-time TM / TE masks, single-row tiling, graph replay, monitors, tangents
    shape2 = (40, 64, 1)
    eps2 = 1 + rng.random(shape2)
    for comp in ("z", "x"):
        F = ceviche_b200.fdtd(eps2, 5e-8, [5, 6, 0], dtype=dtype)
        F.set_option("use_graph", 1)
        n2 = 120
        s2 = [(comp, cases.one_hot(shape2, (20, 30, 0)) + rng.random(shape2) * (rng.random(shape2) < 0.05), cases.gaussian(n2, 30, 8))]
        p2 = [("Ez", rng.random(shape2)), ("Hx", rng.random(shape2)), ("Ex", rng.random(shape2)), ("Hz", rng.random(shape2))]
        ser = F.run(n2, s2, p2, monitors=[("Ez", cases.one_hot(shape2, (10, 12, 0)) + cases.one_hot(shape2, (30, 40, 0)))], freqs=[1e14, 2e14])
        ser = F.run(n2, s2, p2)                                          # graph replay (monitors detached -> graphs on)
        O2 = OracleFDTD(eps2, 5e-8, [5, 6, 0]); o2, _ = O2.run(n2, s2, p2); o3, _ = O2.run(n2, s2, p2)
        print(dtype, comp, "2-D run: series rel-L2", rel_l2(ser.cpu().numpy(), o3), "mask", F._active, flush=True)
        for streams in (0, 1):
            F3 = ceviche_b200.fdtd(eps2, 5e-8, [5, 6, 0], dtype=dtype)
            F3.set_option("jvp_streams", streams)
            _, ds2 = F3.jvp_run(20, torch.as_tensor(rng.standard_normal((3,) + shape2)), [(c, p, w[:20]) for c, p, w in s2], p2)
        print(dtype, comp, "2-D jvp", float(ds2.abs().max()), flush=True)
    F4 = ceviche_b200.fdtd(eps, 5e-8, list(npml), dtype=dtype)           # 3-D marching tangents
    _, ds3 = F4.jvp_run(steps, torch.as_tensor(rng.standard_normal((2,) + shape)), src, prb)
    print(dtype, "3-D jvp", float(ds3.abs().max()), flush=True)
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -15 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -8 gpurun_out/sanitizer_racecheck.log
timeout 600 python -m pytest tests/test_gpu_fields.py -x -q -k edge 2>&1 | tail -3
