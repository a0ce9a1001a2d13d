"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

from conftest import REPO

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--frames", "40"], capture_output=True, text=True, cwd=REPO, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    # the real reference classes when a reference tree is reachable (this container), else the ATen port (the GPU box)
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["batch_per_gpu"] == 32 and "full batch" in d["config"]["note"]       # same config as the native arm
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_native_arm_refuses_to_run_without_cuda():
    """No CPU fallback: the native arm must fail loudly when no CUDA device is visible."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, cwd=REPO, env=env, timeout=600)
    import torch
    if torch.cuda.is_available():   # on the GPU box hiding devices via env may not apply to this interpreter
        return
    assert out.returncode != 0
    assert "no CPU fallback" in (out.stderr + out.stdout)


def test_reference_arm_does_not_load_the_product_library():
    """The reference arm times the reference's CPU path only: it must not import fastvocoder_b200 / map the native .so."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--frames', '20', "
            "'--batch', '2']; runpy.run_path('bench.py', run_name='__main__'); "
            "assert 'fastvocoder_b200' not in sys.modules, 'product package imported'; "
            "assert 'libfastvocoder_b200' not in open('/proc/self/maps').read(), 'native library mapped'")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REPO, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
