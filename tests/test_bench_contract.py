"""bench.py on a box without a GPU: the product arm refuses to run (no CPU fallback), the reference arm
(`--impl reference`: the reference's own PyTorch get_indices from oracle/_ref — the oracle port without it — on the host cores) prints one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                          timeout=600)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour without a CUDA device")
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)


def test_reference_arm_prints_the_contract_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "items/s" and line["higher_is_better"] is True
    assert line["metric"] == "rqvae_semantic_id_encode_items_per_s" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["steps"] == 1 and line["n_gpus"] == 1
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "models", "rqvae.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")      # the real PyTorch reference when it is there
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["cpu_baseline"]["reference_dedup_lane"]["s_per_duplicated_code"] > 0
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
