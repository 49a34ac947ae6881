"""bench.py contract (CPU part): the reference arm runs without a GPU and prints ONE JSON line with the required
keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--layers", "1", "--steps", "1",
                           "--warmup", "0", "--gpus", "1", "--debug-small-cpu"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["workload"].startswith("cfg-2")
    assert d["cfg1"]["cpu_frames_per_s"] > 0          # the same-config pair of the GPU arm's cfg1.gpu_frames_per_s


def test_config_dict_is_the_same_in_every_arm():
    """The driver compares the arms' `config`: it names the workload, never the implementation."""
    sys.path.insert(0, str(ROOT))
    import importlib
    bench = importlib.import_module("bench")
    a = bench.parse.__globals__["argparse"].Namespace(workload="cfg2", layers=42, tiled=False)
    assert bench.config_dict(a, 1) == bench.config_dict(a, 8)


def test_call_timer_work_model_matches_analytic_model():
    """bench.py derives FLOPs from the C-ABI call arguments; summed over the call sequence of one cfg-2 clip they must
    equal the analytic work model (dove_b200.workmodel) that is pinned to the published MAC count."""
    sys.path.insert(0, str(ROOT))
    import bench
    from dove_b200.workmodel import clip_macs
    # conv 128->128 3x3x3 on one frame batch of 8 at 768x1280: 2*27*128*128*8*768*1280
    class P:           # stand-in for ctypes.c_void_p
        def __init__(self, v):
            self.value = v
    ct = bench.CallTimer.__new__(bench.CallTimer)
    ct.cin_real = {1234: 128, 99: 3}
    fl, key, _ = ct._work("dove_conv3d_causal_bf16", (P(1), None, P(1234), None, P(2), 8, 768, 1280, 128, 128, 128))
    assert fl == 2.0 * 27 * 128 * 128 * 8 * 768 * 1280 and key == ("conv", 128, 128, 3, 8, 768, 1280, 1)
    fl, key, _ = ct._work("dove_conv3d_causal_bf16", (P(1), None, P(99), None, P(2), 9, 768, 1280, 64, 128, 128))
    assert fl == 2.0 * 27 * 3 * 128 * 9 * 768 * 1280          # real input channels, not the padded 64
    fl, _, _ = ct._work("dove_attention_bf16", (P(1), P(2), 19426, 48, 0.125))
    assert abs(42 * fl / 2 - clip_macs(33, 768, 1280)["dit_sdpa"]) / clip_macs(33, 768, 1280)["dit_sdpa"] < 1e-12
    fl, _, _ = ct._work("dove_gemm_bf16", (P(1), 3072, P(2), 3072, P(3), 9216, 19426, 9216, 3072))
    assert fl == 2.0 * 19426 * 9216 * 3072


def test_reference_arm_nonzero_rank_is_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
