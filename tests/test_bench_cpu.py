"""bench.py's contract, as far as it can be checked without a GPU: the closed forms the roofline rests on (SURVEY.md
section 8d), the workload bookkeeping (sampler.py:29-35), the reference arm end to end on a tiny sample (one JSON line with
the keys the driver reads; ranks > 0 stay silent), and that the CUDA arm refuses to run without a device."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_work_closed_forms(bench):
    w = bench.algorithmic_work(bench.CONFIGS["C3"])
    assert w["mac"] == 830736 and w["flop"] == 1661472.0              # SURVEY.md 8(d), DESIGN.md section 4
    assert w["tanh"] == 4560 and w["logcosh"] == 676                  # 285 window sites x 16 channels; 13 x 13 x 4
    assert bench.algorithmic_work(bench.CONFIGS["C5"]) == w           # same model: the window does not grow with the lattice
    assert bench.algorithmic_work(bench.CONFIGS["C2"])["flop"] == 85392.0
    c4 = bench.algorithmic_work_total(bench.CONFIGS["C4"])            # 8 images x 2 flipped sites
    assert c4["flop"] == 16 * bench.algorithmic_work(bench.CONFIGS["C4"])["flop"] == 6400.0


def test_workload_bookkeeping(bench):
    its = {k: bench.sample_its(bench.CONFIGS[k]) for k in ("C1", "C2", "C3", "C5")}
    assert its == {"C1": 1441, "C2": 4001, "C3": 16001, "C5": 64001}
    assert [len(bench.flat_params(bench.CONFIGS[k], 0)) for k in ("C1", "C2", "C3")] == [210, 1248, 10600]
    a = bench.config_dict("C3", bench.CONFIGS["C3"], 4096, 1, 0.01)
    assert a["workload"].startswith("C3: 20x20 TFIM") and "4096 chains/GPU" in a["workload"] and a["chains_total"] == 4096
    assert "exceed the 126 MB L2" in a["l2"]                          # 537 MB of caches: inputs larger than L2
    assert "fit the L2" in bench.config_dict("C2", bench.CONFIGS["C2"], 4096, 1, 0.01)["l2"]
    assert abs(bench.cache_mbytes(bench.CONFIGS["C3"], 4096) - 537.0) < 1.0


def test_issue_bound_reads_the_committed_captures_and_never_raises(bench):
    clk = {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0}
    assert bench.issue_bound("C3", 22.4e6, 148, clk) is None          # C3 / C5 are FP32-bound: no issue roofline
    c2 = bench.issue_bound("C2", 177.2e6, 148, clk)                   # the rates of profiles/r02_bench_lines_configs.json
    c4 = bench.issue_bound("C4", 18.3e6, 148, clk)
    assert 4000 < c2["warp_instructions_per_proposal_from_profile"] < 4500 and 0.6 < c2["frac"] < 0.7
    assert 30000 < c4["warp_instructions_per_proposal_from_profile"] < 35000 and 0.45 < c4["frac"] < 0.6
    assert bench.issue_bound("C2", 1.0, 148, None) is None
    assert bench.issue_bound("C2", 1.0, 148, clk, root="/nonexistent") is None


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line_and_only_on_rank_0():
    args = ["--impl", "reference", "--config", "C1", "--steps", "2", "--warmup", "1", "--ref-chains", "8", "--ref-its", "8",
            "--ref-energy", "2"]
    r = _run(args)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "metropolis_proposals_per_s" and d["unit"] == "proposals/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "8 chains x 8" in cb["sample"]
    assert d["config"]["workload"].startswith("C1: 6x6 TFIM")
    # under torchrun only rank 0 runs and prints; the others exit 0 without work
    r1 = _run(args, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r1.returncode == 0 and not [l for l in r1.stdout.splitlines() if l.startswith("{")]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_cuda_arm_refuses_to_run_without_a_device():
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
