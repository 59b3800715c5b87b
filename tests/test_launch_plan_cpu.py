"""Host logic of the sweep planner (qmc_diag_sweep_plan: no CUDA call, runs without a GPU): which persistent kernel
`Sampler.mcmc_op` (sampler.py:158-177) launches for a model / lattice / chain count on a B200 (148 SMs, 227 KB of opt-in
shared memory per CTA), with which geometry.  The BASELINE configurations are pinned to the
geometries the GPU runs of the round recorded (profiles/r02_summary.md); the rest are invariants of the planner."""
import ctypes as C

import pytest

B200_SMS, B200_SMEM = 148, 232448
NONE, W8, W16, W28, IP = -1, 0, 1, 2, 3
CRBM, DCRBM = 0, 1
DEEP = [16, 16, 16, 16, 16, 8]


def plan(kind, k, layers, shape, S, flips, n_steps, reserved=(0, 0, 0, 0), sms=B200_SMS, smem=B200_SMEM):
    from qmcnn_b200 import _lib
    lib = _lib.load()
    d = _lib.ModelDesc()
    d.kind, d.k, d.n_layers = kind, k, len(layers)
    d.Ly, d.Lx = shape
    for i, c in enumerate(layers):
        d.channels[i] = c
    for i, r in enumerate(reserved):
        d.reserved[i] = r
    out = (C.c_int64 * 8)()
    rc = lib.qmc_diag_sweep_plan(C.byref(d), S, flips, n_steps, sms, smem, out)
    assert rc == 0, lib.qmc_last_error(None)
    return dict(zip(("kernel", "ctas", "warps", "smem", "launches", "chunk", "slots"), list(out)[:7]))


def test_baseline_configurations():
    # C1: 64 chains are spread one warp per CTA (not 5 CTAs of 14 warps)
    p = plan(CRBM, 5, [8], (6, 6), 64, 1, 1441)
    assert (p["kernel"], p["ctas"], p["warps"], p["launches"]) == (W8, 64, 1, 1)
    # C2: every layer <= 8 channels -> the 28-warp object, 4096 chains in one wave (ncu: grid 147 x 896 threads, 130.4 kB)
    p = plan(DCRBM, 3, [8, 8, 8], (10, 10), 4096, 1, 4001)
    assert (p["kernel"], p["ctas"], p["warps"], p["smem"], p["launches"]) == (W28, 147, 28, 130432, 1)
    # C3 / C5: the in-place kernel, 12 warps per SM, time-sliced so that every launch is one full wave
    p = plan(DCRBM, 3, DEEP, (20, 20), 4096, 1, 16001)
    assert (p["kernel"], p["ctas"], p["warps"], p["launches"], p["chunk"]) == (IP, 148, 12, 148, 251)
    assert p["smem"] <= B200_SMEM
    p = plan(DCRBM, 3, DEEP, (40, 40), 8192, 1, 64001)
    assert (p["kernel"], p["ctas"], p["warps"], p["launches"], p["chunk"]) == (IP, 148, 12, 296, 1001)
    assert p["smem"] <= B200_SMEM
    # one full wave of chains: a single launch, no slicing
    p = plan(DCRBM, 3, DEEP, (20, 20), 148 * 12, 1, 16001)
    assert (p["kernel"], p["launches"], p["chunk"]) == (IP, 1, 16001)


def test_tuning_knobs_reach_the_planner():
    from qmcnn_b200 import _lib
    # max_warps caps the warps per CTA: C2 at <= 16 warps is the 128-register object in two waves of 14
    p = plan(DCRBM, 3, [8, 8, 8], (10, 10), 4096, 1, 4001, reserved=(0, 16, 0, 0))
    assert (p["kernel"], p["ctas"], p["warps"]) == (W16, 148, 14)
    p = plan(DCRBM, 3, [8, 8, 8], (10, 10), 4096, 1, 4001, reserved=(_lib.FLAG_SWEEP_CLASSIC, 8, 0, 0))
    assert p["kernel"] == W8 and p["warps"] == 7           # 4 waves of 148 x 7 = 4144 slots
    # at <= 8 warps the in-place kernel fits one warp more than the classic one and is chosen
    p = plan(DCRBM, 3, [8, 8, 8], (10, 10), 4096, 1, 4001, reserved=(0, 8, 0, 0))
    assert p["kernel"] == IP and p["warps"] == 8
    # the classic kernel on request, the in-place kernel on request (C2's shape is inside its coverage)
    p = plan(DCRBM, 3, DEEP, (20, 20), 4096, 1, 16001, reserved=(_lib.FLAG_SWEEP_CLASSIC, 0, 0, 0))
    assert p["kernel"] == W8 and p["launches"] == 1
    p = plan(DCRBM, 3, [8, 8, 8], (10, 10), 4096, 1, 4001, reserved=(_lib.FLAG_SWEEP_INPLACE, 0, 0, 0))
    assert p["kernel"] == IP
    # ip_chunks: at most that many chunks per chain
    p = plan(DCRBM, 3, DEEP, (20, 20), 4096, 1, 16001, reserved=(0, 0, 0, 16))
    assert p["kernel"] == IP and p["chunk"] == 1001 and p["launches"] == (16 * 4096 + 1775) // 1776


def test_two_flip_coverage():
    # a CRBM clamps its window to the lattice: always covered
    p = plan(CRBM, 5, [8], (10, 10), 8192, 2, 4001)
    assert p["kernel"] in (W8, W16, W28) and p["slots"] * 2 >= 8192
    # a deep model whose flip box + r - 1 exceeds the lattice is outside the incremental kernels (-> qmc_nd_sweep)
    assert plan(DCRBM, 3, [8, 8, 8], (10, 10), 100, 2, 4001)["kernel"] == NONE
    # ... and inside them on a lattice that is wide enough: box L/2 + 1 = 11, r - 1 = 6 -> 17 <= 20
    assert plan(DCRBM, 3, [8, 8, 8], (20, 20), 100, 2, 4001)["kernel"] in (W8, W16, W28)


@pytest.mark.parametrize("S", [1, 7, 64, 147, 148, 149, 1000, 2072, 2073, 4096, 5000, 8192, 32768])
@pytest.mark.parametrize("model", [(CRBM, 5, [8]), (DCRBM, 3, [8, 8, 8]), (DCRBM, 3, [16, 16, 8]), (DCRBM, 3, DEEP)])
def test_planner_invariants(model, S):
    kind, k, layers = model
    p = plan(kind, k, layers, (20, 20), S, 1, 4001)
    assert p["kernel"] != NONE
    assert 1 <= p["ctas"] <= B200_SMS and p["smem"] <= B200_SMEM
    assert p["slots"] == p["ctas"] * p["warps"]
    limit = {W8: 8, W16: 16, W28: 28, IP: 12}[p["kernel"]]
    assert 1 <= p["warps"] <= limit
    if p["kernel"] == W28:
        assert max(layers) <= 8                       # the 72-register object only for narrow models
    if S <= B200_SMS:
        assert (p["ctas"], p["warps"]) == (S, 1)      # spread: one warp per SM
    if p["kernel"] == IP:
        # time slicing: every launch but the last is one full wave, and the tasks cover every step of every chain
        chunks = -(-4001 // p["chunk"])
        assert p["launches"] == -(-chunks * S // p["slots"])
        assert S <= p["slots"] or p["chunk"] >= 62
    else:
        # persistent: the idle tail of the last wave is small once there is more than a wave of chains
        waves = -(-S // p["slots"])
        if S >= 2 * B200_SMS * limit:
            assert S / (waves * p["slots"]) > 0.8


def test_bad_arguments():
    from qmcnn_b200 import _lib
    lib = _lib.load()
    d = _lib.ModelDesc()
    d.kind, d.k, d.n_layers, d.Ly, d.Lx = CRBM, 4, 1, 6, 6          # even filter
    d.channels[0] = 8
    out = (C.c_int64 * 8)()
    assert lib.qmc_diag_sweep_plan(C.byref(d), 64, 1, 100, B200_SMS, B200_SMEM, out) == -1
    d.k = 5
    assert lib.qmc_diag_sweep_plan(C.byref(d), 0, 1, 100, B200_SMS, B200_SMEM, out) == -1
    assert lib.qmc_diag_sweep_plan(C.byref(d), 64, 3, 100, B200_SMS, B200_SMEM, out) < 0
    # a model whose parameter block does not fit the shared memory of the device
    d.kind, d.k, d.n_layers, d.Ly, d.Lx = DCRBM, 3, 3, 20, 20
    d.channels[0], d.channels[1], d.channels[2] = 128, 128, 8
    assert lib.qmc_diag_sweep_plan(C.byref(d), 64, 1, 100, B200_SMS, B200_SMEM, out) < 0
