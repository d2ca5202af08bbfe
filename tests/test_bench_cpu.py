"""CPU-side contract checks of bench.py and of the host-side dispatch predicates (no GPU)."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` = the reference's own modules (baseline/_ref or /root/reference) on the host
    cores, stepped by the restated worker.py schedule; K timed half-iterations after W warm-up ones."""
    env = dict(os.environ, LCGAN_BENCH_RES="32")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    from oracle import reference_arm as RA
    want = "reference" if RA.find_checkout() else "port"
    assert cb["kind"] == want and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert ("unmodified reference modules" if want == "reference" else "oracle port") in cb["sample"]
    assert d["steps"] == 2 and d["warmup"] == 1 and d["ms_per_step"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_non_zero_ranks():
    env = dict(os.environ, LCGAN_BENCH_RES="32", RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_fused_flow_layer_path_is_selected_for_the_x2_thin_layers_only():
    """Host-side predicate of the all-phase x2 thin kernels (the flow layers, C -> 2)."""
    import torch
    from lcgan_b200 import _lib, ops, plans

    class T:                                  # stride/shape carrier (no device memory needed)
        def __init__(self, shape, strides, dt):
            self.shape, self._st, self.dtype = shape, strides, dt

        def stride(self):
            return self._st

    lib = _lib.lib()
    N, Cin, R = 4, 64, 32
    x = T((N, Cin, R, R), (Cin * R * R, 1, R * Cin, Cin), torch.bfloat16)             # channels-last
    w2 = T((2, 9 * Cin), (9 * Cin, 1), torch.bfloat16)
    plan = plans.conv_transpose_up2(3, R, R)
    d = ops.TapConvDesc()
    y = T((N, 2, 2 * R, 2 * R), (8 * R * R, 1, 4 * R, 2), torch.float32)
    ops._fill_desc(d, plan.launches[0], x, y, Cin, 2, w2, 1.0, 1.0, 1.0, 1.0)
    assert lib.lcgan_tapconv_up2_thin_eligible(C.byref(d)) == 1
    y8 = T((N, 8, 2 * R, 2 * R), (32 * R * R, 1, 16 * R, 8), torch.float32)         # 8 output channels: not thin
    ops._fill_desc(d, plan.launches[0], x, y8, Cin, 8, w2, 1.0, 1.0, 1.0, 1.0)
    assert lib.lcgan_tapconv_up2_thin_eligible(C.byref(d)) == 0
    p1 = plans.conv(3, 1, R, R)                                                      # stride-1 conv: not an x2 layer
    y1 = T((N, 2, R, R), (2 * R * R, 1, 2 * R, 2), torch.float32)
    ops._fill_desc(d, p1.launches[0], x, y1, Cin, 2, w2, 1.0, 1.0, 1.0, 1.0)
    assert lib.lcgan_tapconv_up2_thin_eligible(C.byref(d)) == 0


def test_roofline_summary_of_an_instrumented_cycle():
    sys.path.insert(0, ROOT)
    import bench
    pk = {"hbm": 6500.0, "tensor_burst": 1600.0, "tensor_sustained": 1400.0, "src": "measured"}
    stats = {"tapconv_tc": {"n": 10, "ms": 20.0, "flops": 1.0e13, "bytes": 4.0e10},
             "act_bwd": {"n": 5, "ms": 10.0, "flops": 0, "bytes": 5.0e10},
             "skinny": {"n": 3, "ms": 0.0, "flops": 0, "bytes": 0}}
    roof, rows = bench.summarise_kernels(stats, pk, {"tapconv_tc": 4.2e9})
    assert roof["kernel"] == "tapconv_tc" and roof["bound"] == "tensor" and roof["traffic"] == 4.2e9
    assert abs(roof["achieved"] - 500.0) < 1e-6 and abs(roof["frac"] - 500.0 / 1400.0) < 1e-9
    assert roof["hbm_top"]["kernel"] == "act_bwd" and abs(roof["hbm_top"]["achieved"] - 5000.0) < 1e-6
    assert [r["kernel"] for r in rows] == ["tapconv_tc", "act_bwd", "skinny"]
    assert abs(sum(r["share"] for r in rows) - 1.0) < 1e-9
    # a memory-bound kernel on top
    stats["act_bwd"]["ms"] = 40.0
    roof, _ = bench.summarise_kernels(stats, pk, {})
    assert roof["kernel"] == "act_bwd" and roof["bound"] == "hbm" and roof["traffic"] is None
    assert bench.summarise_kernels({}, pk, {}) == (None, [])
    json.dumps(roof)
    # shape-tagged tap convs: the dominant SHAPE of the dominant kernel is reported, traffic looked up for that shape;
    # a shape below the ridge (AI = 100 FLOP/B < 1400e12 / 6500e9 = 215) is held against the HBM peak
    stats = {"tapconv_tc|A": {"n": 4, "ms": 30.0, "flops": 3.0e12, "bytes": 3.0e10},
             "tapconv_tc|B": {"n": 6, "ms": 10.0, "flops": 1.0e13, "bytes": 1.0e10},
             "box3": {"n": 2, "ms": 5.0, "flops": 0, "bytes": 2.0e10}}
    roof, rows = bench.summarise_kernels(stats, pk, {"tapconv_tc|A": 7.7e9})
    assert roof["kernel"] == "tapconv_tc" and roof["shape"] == "A" and roof["bound"] == "hbm" and roof["traffic"] == 7.7e9
    assert abs(roof["achieved"] - 1000.0) < 1e-6 and abs(roof["bytes_per_launch"] - 7.5e9) < 1
    assert abs(roof["all_shapes"]["tflops"] - 1.3e13 / 0.04 / 1e12) < 1e-6 and len(roof["shapes"]) == 2
    assert rows[0]["kernel"] == "tapconv_tc" and rows[0]["launches"] == 10
    json.dumps(roof)
