"""The reference arm of bench.py (`--impl reference`: the reference's CPU path timed on the host cores)
runs without a GPU, so its JSON contract is checked here: the keys the driver reads, the bounded
sample, and that under torchrun only rank 0 works and prints."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
        "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                           "--steps", "1", "--warmup", "0", "--cpu-events", "12", *flags],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


@pytest.mark.timeout(900)
def test_reference_arm_prints_one_contract_line():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["metric"] == "train rays/s (fwd+bwd)" and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 1
    # the default workload is the north-star configuration: pixel-bandwidth model on
    assert line["config"]["workload"] == "synthetic_pb_on" and line["config"]["bounded_sample"] is True
    assert line["config"]["pixel_bandwidth"] is True and line["config"]["it_sample_size"] == 30
    base = line["cpu_baseline"]
    assert base["kind"] in ("port", "reference") and base["cores"] >= 1 and base["sample"]
    assert base["value"] == line["value"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


@pytest.mark.timeout(300)
def test_reference_arm_other_ranks_exit_silently():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_product_ops_refuse_cpu_tensors():
    """The product path has no CPU / PyTorch fallback: the tensor wrappers of the C ABI raise on CPU
    tensors instead of computing anything (upstream raises NotImplementedError on CPU
    tensors too), the optimizer refuses CPU parameters, and the bench's own arm asserts a device."""
    import torch
    sys.path.insert(0, ROOT)
    from deblur_e_nerf_b200 import ops, optim
    x = torch.zeros(4, 3)
    idx = torch.zeros(4, dtype=torch.int32)
    t = torch.zeros(4)
    offsets = torch.tensor([0, 4], dtype=torch.int32)
    with pytest.raises(NotImplementedError):
        ops.composite(t, torch.zeros(4, 1), t, t, offsets, None)
    with pytest.raises(NotImplementedError):
        ops.segment_sum(x, offsets)
    with pytest.raises(NotImplementedError):
        opt = optim.FusedAdam([torch.nn.Parameter(torch.zeros(3))], lr=0.01)
        opt.param_groups[0]["params"][0].grad = torch.ones(3)
        opt.step()
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0",
                          "--no-cpu-baseline"], capture_output=True, text=True, cwd=ROOT, timeout=300,
                         env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert res.returncode != 0 and "needs a GPU" in res.stderr
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]
