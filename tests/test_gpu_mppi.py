"""GPU tests of the MPPI path (include/odg_mppi.h, opendog_b200/mppi.py): sampling against a numpy restatement of the
Philox stream, the on-device softmin reduction against numpy (float64), and the planner's costs against the same
action sequences stepped through the ordinary batched environment."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_reduce_matches_numpy():
    import ctypes as C
    from opendog_b200 import lib
    L = lib.load()
    T, N, A, lam = 7, 1000, 8, 0.7
    g = torch.Generator().manual_seed(0)
    cost = torch.randn(N, generator=g) * 3 + 5
    act = torch.rand(T, N, A, generator=g) * 2 - 1
    dc, da = cost.cuda(), act.cuda()
    out = torch.empty(T, A, device="cuda"); stats = torch.empty(4, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    lib.check(L.odg_mppi_reduce(p(dc), p(da), T, N, A, lam, p(out), p(stats), None), "reduce")
    c = cost.double().numpy(); w = np.exp(-(c - c.min()) / lam)
    ref = np.einsum("n,tna->ta", w, act.double().numpy()) / w.sum()
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-5
    s = stats.cpu().numpy()
    assert abs(s[0] - c.min()) < 1e-6 and abs(s[1] - c.mean()) < 1e-4 and abs(s[2] - w.sum()) < 1e-3 * w.sum() and int(s[3]) == int(c.argmin())


def test_sample_matches_philox_restatement():
    import ctypes as C
    from opendog_b200 import lib
    from oracle.oracle import philox
    L = lib.load()
    N, A = 64, 8
    mean = torch.linspace(-0.5, 0.5, A).cuda(); act = torch.empty(N, A, device="cuda")
    lib.check(L.odg_mppi_sample(C.c_void_p(mean.data_ptr()), 0.3, N, A, C.c_uint64(99), 5, 17, C.c_void_p(act.data_ptr()), None), "sample")
    a = act.cpu().numpy(); mu = mean.cpu().numpy()
    for n in (0, 13, 63):
        eps = []
        for blk in range(2):
            r = philox(99, n, 5, (17 << 8) | blk, 0x4d505049)
            for pr in range(2):
                u1 = (np.float32(r[2 * pr] >> 8) + np.float32(1.0)) * np.float32(2.0 ** -24)
                u2 = np.float32(r[2 * pr + 1] >> 8) * np.float32(2.0 ** -24)
                rad = math.sqrt(-2.0 * math.log(u1))
                eps += [rad * math.cos(2 * math.pi * u2), rad * math.sin(2 * math.pi * u2)]
        assert np.abs(a[n] - np.clip(mu + 0.3 * np.array(eps), -1, 1)).max() < 1e-5


def test_plan_costs_match_env_rollout_and_graph_replay():
    from opendog_b200.env import BatchedWalkEnv
    from opendog_b200.mppi import MPPI
    N, T = 128, 10
    # a landed robot as the shared start state
    e0 = BatchedWalkEnv(1, seed=3, info_keys=None)
    e0.reset()
    for _ in range(12):
        e0.step(torch.zeros(1, 8, device="cuda"))
    q, v = e0.get_state()
    st = e0.get_env_state()
    m = MPPI(N, T, sigma=0.3, lam=1.0, seed=4, use_graph=True)
    m.set_start(q[0], v[0], last_action=st["last_action"][0], desired_velocity=st["desired_velocity"][0])
    m.plan(update_mean=False)
    torch.cuda.synchronize()
    cost, acts, new_mean = m.cost.clone(), m.actions.clone(), m.new_mean.clone()
    # the same action sequences through an ordinary batched env from the same state
    e = BatchedWalkEnv(N, seed=4, info_keys=("reward_unclipped",), auto_reset=0)
    e.set_state(q.expand(N, -1).contiguous(), v.expand(N, -1).contiguous())
    e.set_env_state(step=torch.zeros(N, dtype=torch.int32), gait_index=torch.zeros(N, dtype=torch.int32),
                    gait_matches=torch.zeros(N, dtype=torch.int32), last_action=st["last_action"].expand(N, -1).contiguous(),
                    desired_velocity=st["desired_velocity"].expand(N, -1).contiguous(), fresh=torch.zeros(N, dtype=torch.uint8))
    ref = torch.zeros(N, device="cuda"); alive = torch.ones(N, dtype=torch.bool, device="cuda")
    for t in range(T):
        _, _, _, info = e.step(acts[t])
        r = info["reward_unclipped"]
        term = e.terminated.bool()
        stepped = torch.where(alive, ref - r, ref)                       # (same float32 order as the rollout kernel)
        ref = torch.where(alive & term, stepped + 100.0, stepped)
        alive = alive & ~term
    assert torch.equal(ref, cost) and cost.abs().min().item() > 0, "fused rollout != the same sequences stepped one launch at a time"
    # softmin mean of the sampled sequences
    w = torch.exp(-(cost - cost.min()) / 1.0).double()
    refm = torch.einsum("n,tna->ta", w, acts.double()) / w.sum()
    assert (new_mean.double() - refm).abs().max().item() < 2e-5
    # second call captures the graph, third replays it. The plan counter lives in device memory and is bumped inside the
    # graph, so every replay draws fresh noise — the rows of plan k are the stand-alone sampler's with iteration = k.
    import ctypes as C
    from opendog_b200 import lib
    m.plan(update_mean=False); torch.cuda.synchronize(); c2, a2 = m.cost.clone(), m.actions.clone()
    m.plan(update_mean=False); torch.cuda.synchronize(); c3, a3 = m.cost.clone(), m.actions.clone()
    assert m.graph is not None and int(m.iteration_dev) == 3
    assert not torch.equal(c2, cost) and not torch.equal(c3, c2) and not torch.equal(a3, a2)
    row = torch.empty(N, 8, device="cuda")
    for it, acts_it in ((1, a2), (2, a3)):
        lib.check(m.L.odg_mppi_sample(C.c_void_p(m.mean[5].data_ptr()), 0.3, N, 8, C.c_uint64(4), it, 5, C.c_void_p(row.data_ptr()), None), "sample")
        assert torch.equal(row, acts_it[5])


def test_reduce_accepts_the_documented_sample_limit():
    """odg_mppi_reduce keeps [n_samples] weights in dynamic shared memory: 12000 samples need 48 KB on top of 8 KB of
    static rows, i.e. the opt-in limit."""
    import ctypes as C
    from opendog_b200 import lib
    L = lib.load()
    T, N, A = 3, 12000, 8
    cost = torch.rand(N, device="cuda"); act = torch.rand(T, N, A, device="cuda")
    out = torch.empty(T, A, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    lib.check(L.odg_mppi_reduce(p(cost), p(act), T, N, A, 1.0, p(out), None, None), "reduce")
    torch.cuda.synchronize()
    w = torch.exp(-(cost - cost.min())).double()
    ref = torch.einsum("n,tna->ta", w, act.double()) / w.sum()
    assert (out.double() - ref).abs().max().item() < 2e-5
