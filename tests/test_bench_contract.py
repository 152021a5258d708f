"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm runs the reference's
own hot path on the host (oracle/ref_shim.py: source tree or the byte code in oracle/_ref; the oracle port only when
neither exists) and prints ONE JSON line with the agreed keys; non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *args):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    out = _run(None, "--impl", "reference", "--steps", "1", "--warmup", "0", "--seq-len", "128")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "docs_per_sec_heads_plus_decode" and d["unit"] == "docs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1
    import ref_shim

    want_kind = "reference" if ref_shim.reference_available() else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_non_zero_rank_is_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--steps", "1", "--warmup", "0",
               "--seq-len", "64", "--gpus", "2")
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_benchmark_scripts_parse_and_pipeline_defaults():
    """Every script under benchmarks/ is valid Python; the serving loop keeps three batches in flight and offers the
    GC freeze the end-to-end leg of bench.py relies on."""
    import ast
    import glob
    import inspect

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scripts = sorted(glob.glob(os.path.join(root, "benchmarks", "*.py")))
    assert scripts, "benchmarks/ is empty"
    for path in scripts:
        with open(path) as f:
            ast.parse(f.read(), filename=path)
    from peneo_b200 import HeadsDecodePipeline

    assert inspect.signature(HeadsDecodePipeline.run).parameters["depth"].default == 3
    assert callable(HeadsDecodePipeline.freeze_host_gc)
