"""bench.py contract: ONE JSON line on stdout with the keys the driver reads.  The CUDA arm needs a
GPU (covered by -m gpu); the reference arm (CPU port of the reference path) runs anywhere."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout=600):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must be exactly one line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--nside", "32"])
    assert COMMON <= set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "evals/s" and d["value"] > 1e6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "precision_mode" not in d["config"]
    # one step = one pass over the SAMPLE: the measured time is reported, the full-map figure is extrapolated
    assert d["ms_per_full_map_extrapolated"] >= d["ms_per_step"] > 0


@pytest.mark.gpu
def test_cuda_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--nside", "256", "--no-cpu-baseline", "--no-configs"])
    assert COMMON | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["gpu_launches"] == 3 and d["value"] > 1e10
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["kernel"].startswith("zodi_los_")
    # executed-pipe utilisations come from the committed ncu counts (profiles/kernel_counts.json)
    assert r["executed"] and 0.3 < r["frac_executed"] < 1.1 and r["limiter"]["pipe"] in ("issue_slot", "xu_pipe", "fma_pipe")
    assert d["max_rel_err_vs_oracle"] <= d["tolerance"]
    assert d["clocks"]["samples"] >= 20
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] > 0
    assert 0.2 < e["link_frac"] < 1.2
    assert e["healpix_entry"]["max_rel_diff_vs_array_seam"] == 0.0
    assert d["fp64_mode"]["e2e"]["value"] > 0 and d["fp64_mode"]["max_rel_err_vs_oracle"] <= 1e-10


@pytest.mark.gpu
def test_cuda_arm_configs_block():
    """The other BASELINE configurations ride on the same line: 1, 2, 4 (time-ordered data through the
    on-device ephemeris, observer = semb-l2; reduced to 1e6 samples here) and 5, fp32 + fp64, each within
    its tolerance of the oracle."""
    d = _run(["--steps", "2", "--warmup", "3", "--nside", "128", "--no-cpu-baseline", "--no-e2e", "--tod-samples",
              "1e6"], timeout=900)
    cfg = d["configs"]
    assert set(cfg) == {"1", "2", "4", "5"}
    for key, c in cfg.items():
        for precision in ("fp32", "fp64"):
            assert c[precision]["ok"], (key, precision, c[precision])
            assert c[precision]["value"] > 1e10
    assert cfg["4"]["n_los"] == 1_000_000 and set(cfg["4"]["e2e"]) == {"unit_vectors_32B_per_sample", "lonlat_24B_per_sample"}
    assert abs(cfg["4"]["semb_l2_scale"]["host"] / cfg["4"]["semb_l2_scale"]["device"] - 1.0) < 1e-14


def test_reference_and_cuda_arms_share_one_config():
    """The driver compares the `config` dicts of the two arms: they must be produced by one function."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.workload_config(2048) == bench.workload_config(2048)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": workload_config(args.nside)') == 2
