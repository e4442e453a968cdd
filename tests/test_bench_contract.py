"""bench.py contract checks that need no GPU: the reference arm's JSON line and the static structure of our arm's line."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_json_line():
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "cartpole_rn"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "se_env_steps_per_s" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    # the unmodified reference when it is staged (oracle/make_ref.py -> oracle/_ref/pyref) with the C port as a second figure,
    # else the port alone
    assert cb["kind"] in ("reference", "port") and 1 <= cb["cores"] <= (os.cpu_count() or 1) and cb["value"] == d["value"] and "sample" in cb
    if cb["kind"] == "reference":
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 0 and d["nes_generations_per_hour"] > 0 and d["nes_population"] == 16
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_nonzero_ranks():
    env = dict(os.environ, PYTHONPATH=ROOT, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_our_arm_line_carries_every_contract_key():
    """Static check of bench.py's source: the keys of the JSON line our arm prints (needs a GPU to run)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def measure_nes("):]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert re.search(r'"%s":' % key, body) or re.search(r'res\["%s"\] =' % key, body), key
    assert 'line["cpu_baseline"] =' in body
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert re.search(r'"%s":' % key, body[body.index('"roofline"'):]), key
    assert "no CPU fallback" in src           # our arm refuses to run without a CUDA device
