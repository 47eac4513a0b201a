"""bench.py's reference arm (`--impl reference`) runs without a GPU: check the JSON line the driver parses.

The arm times the reference's own kernel (oracle/_ref, or the oracle port when that artefact is missing) on the host
cores; here only the shape of the line is checked, on the smallest sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, args=()):
    env = dict(os.environ, PTB200_BENCH_CPU_SAMPLE="small")
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", *args],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["unit"] == "Mpaths/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "c2" and d["config"]["paths_per_gpu"] == 1024 * 768 * 64
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "paths" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, args=("--gpus", "2"))
    assert out.strip() == ""


def test_tools_and_bench_compile():
    """The measurement scripts are run by hand on the GPU box: at least keep them syntactically alive."""
    import glob
    import py_compile
    for path in sorted(glob.glob(os.path.join(ROOT, "tools", "*.py"))) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        py_compile.compile(path, doraise=True)
