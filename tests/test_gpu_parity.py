"""GPU parity: the CUDA path (through the C ABI) against the fp64 oracle on identical inputs.

Stated fp32 tolerances, single mj_step from identical state:
    |dqpos| <= 1e-6 + 1e-5*|qpos|,  |dqvel| <= 1e-4 + 1e-3*|qvel|,  contact normal force rel 1e-2.
Reward / obs within 1e-5; done, truncation, gait integers and reset indexing bit-exact.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_emu_parity import FLIP_M, assert_info_matches  # noqa: E402  (shared tolerances; importing runs nothing)

# Launch shapes of the step kernel (OdgEnvConfig::launch_*). "auto" is what a small handle gets; "large-batch" is what every
# batch of more than one wave gets (>= ~8192 envs: BASELINE configs[2]/[3]); the others are the remaining code paths
# (2 / 4 environments per warp, one- and four-warp blocks). Every oracle comparison below runs under each of them.
SHAPES = {
    "auto": {},
    "large-batch": dict(launch_lanes=32, launch_block=128, launch_lockstep=2, launch_fat=1),  # what every batch of more than one wave gets
    "lockstep": dict(launch_lanes=32, launch_block=64, launch_lockstep=1, launch_fat=0),      # lean instantiation, whole-block lockstep
    "lean": dict(launch_fat=0),
    "2-per-warp": dict(launch_lanes=8, launch_block=128, launch_lockstep=0),
    "4-per-warp-lockstep": dict(launch_lanes=16, launch_block=32, launch_lockstep=1),
}
shapes = pytest.mark.parametrize("shape", list(SHAPES), ids=list(SHAPES))
GPU_INFO = dict(paw="paw_contact_forces", lin="linear_vel_tracking_reward", dist="distance_from_origin")


def _np_info(info):
    return {k: v.cpu().numpy() for k, v in info.items()}


def _oracle_rollout_states(n_states, seed=0, settle=0):
    """States visited by the oracle under random ctrl targets (mix of flight, impact and stance)."""
    from oracle.oracle import Sim
    rng = np.random.default_rng(seed)
    s = Sim()
    s.reset_keyframe()
    lo = np.array([2.36, -1.8] * 4); hi = np.array([2.8, -1.2] * 4)
    out = []
    k = 0
    while len(out) < n_states:
        if k % 25 == 0:
            s.ctrl[:] = lo + rng.uniform(0, 1, 8) * (hi - lo)
        s.step()
        k += 1
        if k > settle and k % 7 == 0:
            out.append((s.qpos.copy(), s.qvel.copy(), s.qacc_warmstart.copy(), s.ctrl.copy()))
        if k % 900 == 0:
            s.reset_keyframe()
            s.qpos[:3] += rng.uniform(-0.02, 0.02, 3)
    return out


@shapes
def test_single_step_physics_parity(shape):
    from opendog_b200.env import BatchedWalkEnv
    from oracle.oracle import Sim
    states = _oracle_rollout_states(256, seed=1)
    N = len(states)
    env = BatchedWalkEnv(N, frame_skip=1, scale_actions=0, auto_reset=0, solver_iterations=30,
                         info_keys=("qacc", "ncon", "contact_normal_force", "solver_iters"), **SHAPES[shape])
    qpos = np.stack([s[0] for s in states]).astype(np.float32)
    qvel = np.stack([s[1] for s in states]).astype(np.float32)
    warm = np.stack([s[2] for s in states]).astype(np.float32)
    ctrl = np.stack([s[3] for s in states]).astype(np.float32)
    env.set_state(qpos, qvel, warm)
    _, _, _, info = env.step(torch.from_numpy(ctrl).cuda())
    gq, gv = [t.cpu().numpy() for t in env.get_state()]
    ncon = info["ncon"].cpu().numpy(); fn = info["contact_normal_force"].cpu().numpy()
    sim = Sim()
    bad = 0
    for i in range(N):
        sim.reset_keyframe()
        sim.qpos[:] = qpos[i].astype(np.float64); sim.qvel[:] = qvel[i].astype(np.float64)
        sim.qacc_warmstart[:] = warm[i].astype(np.float64); sim.ctrl[:] = ctrl[i].astype(np.float64)
        sim.step()
        ofn = sum(c["force"][0] for c in sim.contacts())
        ok = (np.all(np.abs(gq[i] - sim.qpos) <= 1e-6 + 1e-5 * np.abs(sim.qpos)) and
              np.all(np.abs(gv[i] - sim.qvel) <= 1e-4 + 1e-3 * np.abs(sim.qvel)) and
              abs(fn[i] - ofn) <= 1e-2 * max(1.0, abs(ofn)) and ncon[i] == sim.ncon)
        if not ok:
            # only a discrete decision at fp32 resolution may explain it (see FLIP_M), and the oracle must say so
            bad += 1
            g = sim.decision_gaps
            assert min(g[0], g[1]) < FLIP_M, f"state {i}: outside the single-step tolerance without a decision flip {g}"
    assert bad <= max(1, N // 100), f"{bad}/{N} envs outside the stated single-step tolerance"


@shapes
def test_walk_env_parity_and_reset_indexing(shape):
    """Rows a2-a12: 64 envs x 60 steps against the oracle, re-synchronised every step. done / truncation / reset indexing
    bit-exact for every env and step; obs, reward and EVERY info output (x/y position, distance, reward_ctrl,
    linear_vel_tracking_reward, patterns_matches, paw_contact_forces with the quirks of reward_calc.py:339-370,
    terminal_obs) within the stated tolerances; an env-step outside them must be a decision flip the oracle confirms."""
    from opendog_b200.env import BatchedWalkEnv
    from oracle.oracle import WalkEnv
    N, T = 64, 60
    env = BatchedWalkEnv(N, seed=11, max_episode_steps=25, solver_iterations=30,
                         info_keys=("terminal_obs", "gait_reward", "paws_in_ground", "patterns_matches", "x_position",
                                    "y_position", "distance_from_origin", "paw_contact_forces",
                                    "linear_vel_tracking_reward", "reward_ctrl"), **SHAPES[shape])
    ws = [WalkEnv(seed=11, env_id=i) for i in range(N)]
    for w in ws:
        w.e.max_steps = 25
    obs = env.reset().cpu().numpy()
    oobs = np.stack([w.reset() for w in ws])
    gq, gv = [x.cpu().numpy() for x in env.get_state()]
    assert np.array_equal(gq, np.stack([w.qpos for w in ws]).astype(np.float32)), "reset state must be bit-exact"
    assert not gv.any()
    assert np.abs(obs - oobs).max() < 1e-6, "reset obs"
    rng = np.random.default_rng(5)
    n_done = outliers = checked_forces = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        obs, rew, done, info = env.step(torch.from_numpy(a).cuda())
        obs = obs.cpu().numpy(); rew = rew.cpu().numpy(); done = done.cpu().numpy(); info = _np_info(info)
        trunc = env.truncated.cpu().numpy().astype(bool); term = env.terminated.cpu().numpy().astype(bool)
        res = [w.step_autoreset(a[i]) for i, w in enumerate(ws)]
        odone = np.array([r[2] for r in res])
        assert np.array_equal(done, odone), f"step {t}: done / reset indexing differs"
        assert np.array_equal(trunc & ~term, np.array([r[3] for r in res])), f"step {t}: TimeLimit.truncated differs"
        n_done += int(odone.sum())
        for i, (oo, orr, od, _, otob, oi) in enumerate(res):
            # the episode's last observation: terminal_obs when the env auto-reset, else the returned obs
            mine, theirs = (info["terminal_obs"][i], otob) if od else (obs[i], oo)
            ok = (np.abs(mine - theirs).max() < 2e-4 and abs(rew[i] - orr) < 2e-4
                  and info["gait_reward"][i] == oi["gait_first_call"]
                  and np.array_equal(info["paws_in_ground"][i], oi["paws_in_ground"]))
            if ok:
                assert_info_matches(info, oi, i, GPU_INFO)
                checked_forces += int(oi["paws_in_ground"].any())
                if od:
                    assert np.abs(obs[i] - oo).max() < 1e-6, "obs returned for a done env is its reset obs"
            else:
                outliers += 1
                assert min(oi["min_gap"][0], oi["min_gap"][1]) < FLIP_M, \
                    f"step {t} env {i}: outside tolerance without a decision flip (gaps {oi['min_gap']})"
                assert np.abs(mine - theirs).max() < 5e-2, f"step {t} env {i}: gross divergence"
        # chaotic contact dynamics: re-synchronise the oracle to the GPU state every step so the test
        # measures per-step agreement rather than trajectory divergence
        gq, gv = [x.cpu().numpy() for x in env.get_state()]
        for i, w in enumerate(ws):
            w.qpos[:] = gq[i]; w.qvel[:] = gv[i]
    assert outliers <= max(2, (N * T) // 500), f"{outliers}/{N * T} env-steps are decision flips (expected <= 0.2 %)"
    assert n_done >= N, "truncation/auto-reset path was not exercised"
    assert checked_forces > N * T // 3, "paw contact forces were not exercised"


@shapes
def test_evaluate_logic_bit_exact(shape):
    """Reward / termination / gait logic on IDENTICAL states: integers and flags bit-exact, reward to 2 ulp."""
    from opendog_b200.env import BatchedWalkEnv
    from oracle.oracle import WalkEnv
    states = _oracle_rollout_states(128, seed=3, settle=150)
    N = len(states)
    rng = np.random.default_rng(9)
    env = BatchedWalkEnv(N, seed=2, auto_reset=0, solver_iterations=30,
                         info_keys=("gait_reward", "paws_in_ground", "patterns_matches", "linear_vel_tracking_reward",
                                    "x_position", "y_position", "distance_from_origin", "reward_ctrl", "paw_contact_forces"),
                         **SHAPES[shape])
    qpos = np.stack([s[0] for s in states]).astype(np.float32)
    qvel = np.stack([s[1] for s in states]).astype(np.float32)
    # tilt some trunks past / near the 15 degree limit and push some forward to exercise every branch
    for i in range(0, N, 4):
        ang = rng.uniform(0.2, 0.35); ax = rng.integers(1, 4)
        q = np.zeros(4); q[0] = np.cos(ang / 2); q[ax] = np.sin(ang / 2)
        qpos[i, 3:7] = q
    qpos[1::3, 0] = rng.uniform(0.1, 2.0, len(qpos[1::3])); qvel[1::3, 0] = rng.uniform(0.4, 1.2, len(qvel[1::3]))
    ctrl = np.stack([s[3] for s in states]).astype(np.float32)
    last = rng.uniform(-2, 2.8, (N, 8)).astype(np.float32)
    gidx = rng.integers(0, 8, N).astype(np.int32); gcnt = (8 * rng.integers(0, 5, N)).astype(np.int32)
    step = rng.integers(0, 760, N).astype(np.int32); fresh = (rng.uniform(size=N) < 0.2).astype(np.uint8)
    last[fresh.astype(bool)] = 0
    env.set_state(qpos, qvel)
    env.set_env_state(step=step, gait_index=gidx, gait_matches=gcnt, last_action=last, fresh=fresh)
    desvel = env.get_env_state()["desired_velocity"].cpu().numpy()
    obs, rew, term, trunc, info = env.evaluate(torch.from_numpy(ctrl).cuda())
    obs = obs.cpu().numpy(); rew = rew.cpu().numpy(); term = term.cpu().numpy(); trunc = trunc.cpu().numpy()
    ninfo = _np_info(info)
    mism = 0
    for i in range(N):
        w = WalkEnv(seed=2, env_id=i)
        assert np.array_equal(np.float32(w.desired_velocity), desvel[i])
        w.qpos[:] = qpos[i]; w.qvel[:] = qvel[i]
        w.e.step = int(step[i]); w.e.gait_index = int(gidx[i]); w.e.gait_matches = int(gcnt[i])
        w.last_action[:] = last[i]; w.e.last_action_is_reset = int(fresh[i])
        oo, orr, ot, otr, oi = w.evaluate(ctrl[i])
        same_contacts = np.array_equal(oi["paws_in_ground"], ninfo["paws_in_ground"][i])
        if not same_contacts:
            mism += 1
            assert oi["min_gap"][0] < FLIP_M, f"state {i}: paw contact flags differ without a margin flip {oi['min_gap']}"
            continue
        if oi["min_gap"][0] >= FLIP_M and oi["min_gap"][1] >= FLIP_M:
            assert_info_matches(ninfo, oi, i, GPU_INFO)
        # x / y / distance are pure functions of the (identical) state: exact in float32
        assert ninfo["x_position"][i] == qpos[i, 0] and ninfo["y_position"][i] == qpos[i, 1]
        assert np.isclose(ninfo["distance_from_origin"][i], np.hypot(qpos[i, 0], qpos[i, 1]), rtol=1e-6, atol=0)
        assert ot == term[i] and otr == trunc[i]
        assert oi["gait_first_call"] == int(info["gait_reward"][i]) and oi["patterns_matches"] == float(info["patterns_matches"][i])
        assert np.array_equal(oo.astype(np.float32), obs[i]) or np.abs(oo - obs[i]).max() < 1e-6
        assert abs(np.float32(orr) - rew[i]) <= 2 * np.spacing(np.float32(max(abs(orr), 1e-3)))
    assert mism <= 2, "paw contact flags disagree on identical states"
    assert term.any() and (~term).any() and trunc.any()


def test_go1_physics_parity_and_env():
    """Second model descriptor (Unitree Go1, go1.xml: 12 actuators, 3 joints per leg, joint damping -> implicit-damping
    Euler, sphere feet) through the same kernel: single mj_step from identical states vs the oracle at the stated
    tolerance, then the walk task on it (obs 9 + 3*12 = 45, the layout of landing_environment.py:116-136)."""
    from opendog_b200.env import BatchedWalkEnv
    from oracle.oracle import Sim
    rng = np.random.default_rng(4)
    s = Sim("go1"); s.reset_keyframe()
    lo = np.array([r[0] for r in s.desc["act_ctrlrange"]]); hi = np.array([r[1] for r in s.desc["act_ctrlrange"]])
    mid = np.array(s.desc["key_ctrl"])
    states = []
    for k in range(900):
        if k % 40 == 0:
            s.ctrl[:] = np.clip(mid + rng.uniform(-0.15, 0.15, 12), lo, hi)
        s.step()
        if k % 6 == 0:
            states.append((s.qpos.copy(), s.qvel.copy(), s.qacc_warmstart.copy(), s.ctrl.copy()))
    N = len(states)
    env = BatchedWalkEnv(N, model="go1", frame_skip=1, scale_actions=0, auto_reset=0, solver_iterations=100,
                         info_keys=("ncon", "contact_normal_force"))
    f32 = lambda i: np.stack([st[i] for st in states]).astype(np.float32)
    env.set_state(f32(0), f32(1), f32(2))
    _, _, _, info = env.step(torch.from_numpy(f32(3)).cuda())
    gq, gv = [t.cpu().numpy() for t in env.get_state()]
    ncon = info["ncon"].cpu().numpy(); fn = info["contact_normal_force"].cpu().numpy()
    bad = 0
    for i in range(N):
        s.reset_keyframe()
        s.qpos[:] = f32(0)[i]; s.qvel[:] = f32(1)[i]; s.qacc_warmstart[:] = f32(2)[i]; s.ctrl[:] = f32(3)[i]
        s.step()
        ofn = sum(c["force"][0] for c in s.contacts())
        ok = (np.all(np.abs(gq[i] - s.qpos) <= 1e-6 + 1e-5 * np.abs(s.qpos)) and
              np.all(np.abs(gv[i] - s.qvel) <= 1e-4 + 1e-3 * np.abs(s.qvel)) and
              abs(fn[i] - ofn) <= 1e-2 * max(1.0, abs(ofn)) and ncon[i] == s.ncon)
        if not ok:
            bad += 1             # a sphere foot within fp32 rounding of the 1 mm margin: the oracle must confirm it
            assert s.decision_gaps[0] < FLIP_M, f"Go1 state {i}: outside tolerance without a margin flip {s.decision_gaps}"
    assert bad <= max(2, N // 50), f"{bad}/{N} Go1 states outside the single-step tolerance"
    assert ncon.max() >= 4 and ncon.min() >= 1          # four feet; now and then a calf capsule's end as well
    # the env surface on Go1
    e = BatchedWalkEnv(256, model="go1", seed=3, info_keys=None)
    obs = e.reset()
    assert obs.shape == (256, 48) and e.act_dim == 12          # BASELINE configs[2]: policy MLP 48-512-256-12
    e45 = BatchedWalkEnv(256, model="go1", seed=3, info_keys=None, obs_layout=0)
    o45 = e45.reset()
    assert o45.shape == (256, 45)
    # layout 1 = [v, w, projected_gravity, v_des, q - key_qpos, qd, last_action]; the shared entries agree with layout 0
    assert torch.equal(obs[:, :6], o45[:, :6]) and torch.equal(obs[:, 9:12], o45[:, 6:9]) and torch.equal(obs[:, 24:], o45[:, 21:])
    q0, _ = e.get_state()
    from opendog_b200.model.compile import load_compiled
    key = torch.tensor(load_compiled("go1")["key_qpos"][7:19], dtype=torch.float32, device="cuda")
    assert torch.allclose(obs[:, 12:24], q0[:, 7:19] - key, atol=1e-6)
    # projected_gravity is the reference's Euler-angle formula (landing_environment_reward_calc.py:88-98), in double
    qq = q0[:, 3:7].double().cpu().numpy()
    w, x, y, z = qq.T
    eul = np.stack([np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)), np.arcsin(np.clip(2 * (w * y - z * x), -1, 1)),
                    np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))], 1)
    v = (eul @ np.array([0.0, 0.0, -9.81]))[:, None] * eul
    pg = v / np.linalg.norm(v, axis=1, keepdims=True)
    assert np.allclose(obs[:, 6:9].cpu().numpy(), pg, atol=1e-6)
    for t in range(20):
        obs, rew, done, _ = e.step((torch.rand(256, 12, device="cuda") * 2 - 1) * 0.3)
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and (rew >= 0).all()


def test_physics_failure_is_contained_per_environment():
    """Failure detection / recovery: a non-finite state in one environment becomes terminated = 1
    for THAT environment (is_healthy, reward_calc:117-121; the `except mujoco.FatalError` path of
    sim2real/train.py:282-283), the auto-reset brings it back, and its neighbours in the same warp are untouched."""
    from opendog_b200.env import BatchedWalkEnv
    N = 64
    a = torch.rand(6, N, 8, device="cuda") * 2 - 1
    ref = BatchedWalkEnv(N, seed=21, info_keys=None)
    env = BatchedWalkEnv(N, seed=21, info_keys=("terminal_obs",))
    ref.reset(); env.reset()
    for t in range(2):
        ref.step(a[t]); env.step(a[t])
    q, v = env.get_state()
    bad = [3, 17, 18, 40]
    q[bad[0], 2] = float("nan"); v[bad[1], 7] = float("inf"); q[bad[2], 9] = float("nan"); v[bad[3], 0] = float("-inf")
    qr, vr = ref.get_state()
    env.set_state(q, v); ref.set_state(qr, vr)               # (both sides restart their solver warm start)
    obs, rew, done, info = env.step(a[2]); robs, rrew, rdone, _ = ref.step(a[2])
    term = env.terminated.bool().cpu()
    assert term[bad].all() and done.cpu()[bad].all()
    good = torch.ones(N, dtype=torch.bool); good[bad] = False
    assert torch.equal(obs.cpu()[good], robs.cpu()[good]) and torch.equal(rew.cpu()[good], rrew.cpu()[good])
    assert torch.isfinite(obs).all(), "the returned obs of a failed env is its reset obs"
    q2, v2 = env.get_state()
    assert torch.isfinite(q2).all() and torch.isfinite(v2).all()
    for t in range(3, 6):
        obs, rew, done, _ = env.step(a[t])
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()


def test_determinism_sharding_and_soak():
    """(1) Same seed twice -> bit-identical trajectories. (2) A handle of 128 envs equals two handles of 64 with
    first_env_id 0 / 64 (the layout of two ranks): RNG streams are keyed by global env id and an environment's result
    does not depend on which environments share its warp or block. (3) Soak: 1024 envs x 600 random-action steps stay
    finite, episodes end and restart, rewards stay in range."""
    from opendog_b200.env import BatchedWalkEnv
    g = torch.Generator(device="cuda").manual_seed(7)
    acts = torch.rand(25, 128, 8, device="cuda", generator=g) * 2 - 1

    def run(n, first, sl, **shape):
        e = BatchedWalkEnv(n, seed=5, first_env_id=first, info_keys=None, max_episode_steps=12, **shape)
        out = [e.reset().clone()]
        for t in range(25):
            o, r, d, _ = e.step(acts[t, sl].contiguous())
            out += [o.clone(), r.clone(), d.clone()]
        return out
    a, b = run(128, 0, slice(0, 128)), run(128, 0, slice(0, 128))
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # the launch shape is a host-side choice (2, 4 or 8 environments per warp; 1-4 warps per block; Newton iterations
    # free-running or in block lockstep, which is what large batches get): it must not change a single bit
    for name, shape in SHAPES.items():
        c = run(128, 0, slice(0, 128), **shape)
        assert all(torch.equal(x, y) for x, y in zip(a, c)), f"launch shape {name} changed a result"
    lo, hi = run(64, 0, slice(0, 64)), run(64, 64, slice(64, 128))
    for x, y, z in zip(a, lo, hi):
        assert torch.equal(x, torch.cat([y, z]))
    e = BatchedWalkEnv(1024, seed=9, info_keys=None, max_episode_steps=100)
    e.reset()
    ndone, rmax = 0, 0.0
    for t in range(600):
        o, r, d, _ = e.step(torch.rand(1024, 8, device="cuda") * 2 - 1)
        ndone += int(d.sum()); rmax = max(rmax, float(r.max()))
    q, v = e.get_state()
    assert torch.isfinite(o).all() and torch.isfinite(q).all() and torch.isfinite(v).all()
    assert ndone >= 5 * 1024 and 0.0 <= rmax < 200.0
    assert float(q[:, 2].min()) > -0.01 and float(q[:, 2].max()) < 0.5          # nobody fell through the floor or flew away


def test_full_size_batches_equal_small_batches_and_hold_their_weight():
    """BASELINE.json's full sizes through size-independent properties. (1) The first and last 64 environments of a
    65536-environment handle (configs[3]) are bit-identical to 64-environment handles with the same global ids, each
    handle with the launch shape its own size selects (the big ones iterate in block lockstep, the small ones do not).
    (2) Same for 16384 (configs[2]) and 4096 environments (configs[1]). (3) 4096 robots holding the home pose come to
    rest carrying their weight: sum of contact normal forces = total mass x 9.81 within 2 %, quaternions stay
    normalised, nobody sinks below the floor."""
    from opendog_b200.env import BatchedWalkEnv
    from opendog_b200.model.compile import load_compiled
    for n in (65536, 16384, 4096):
        g = torch.Generator(device="cuda").manual_seed(n)
        acts = torch.rand(6, n, 8, device="cuda", generator=g) * 2 - 1
        big = BatchedWalkEnv(n, seed=13, info_keys=None)
        lo = BatchedWalkEnv(64, seed=13, info_keys=None, launch_fat=(0 if n == 16384 else -1))     # (one of the small ones: the lean instantiation)
        hi = BatchedWalkEnv(64, seed=13, first_env_id=n - 64, info_keys=None)
        ob, ol, oh = big.reset(), lo.reset(), hi.reset()
        assert torch.equal(ob[:64], ol) and torch.equal(ob[-64:], oh)
        for t in range(6):
            ob, rb, db, _ = big.step(acts[t])
            ol, rl, dl, _ = lo.step(acts[t, :64].contiguous())
            oh, rh, dh, _ = hi.step(acts[t, -64:].contiguous())
            assert torch.equal(ob[:64], ol) and torch.equal(rb[:64], rl) and torch.equal(db[:64], dl), (n, t)
            assert torch.equal(ob[-64:], oh) and torch.equal(rb[-64:], rh) and torch.equal(db[-64:], dh), (n, t)
        assert torch.isfinite(ob).all() and torch.isfinite(rb).all()
        del big, lo, hi
    desc = load_compiled("our_robot")
    weight = (desc["base_mass"] + float(np.sum(desc["mass"]))) * 9.81
    env = BatchedWalkEnv(4096, seed=2, info_keys=("contact_normal_force",), scale_actions=0, max_episode_steps=10**6)
    env.reset()
    home = torch.tensor(desc["key_ctrl"], dtype=torch.float32, device="cuda").expand(4096, 8).contiguous()
    for t in range(60):
        obs, rew, done, info = env.step(home)
    q, v = env.get_state()
    rest = (v.abs().max(dim=1).values < 0.05) & ~done
    assert int(rest.sum()) > 3500, int(rest.sum())
    fn = info["contact_normal_force"][rest]
    assert float((fn / weight - 1).abs().max()) < 0.02, (float(fn.min()), float(fn.max()), weight)
    assert float((q[:, 3:7].norm(dim=1) - 1).abs().max()) < 1e-3 and float(q[:, 2].min()) > 0.0


@pytest.mark.parametrize("n", [16384, 65536])
def test_large_handles_match_the_oracle_on_random_subsets(n):
    """BASELINE configs[2] / [3] batch sizes against the ORACLE (not against smaller handles): run the full handle —
    whatever launch shape its size selects, i.e. block-lockstep Newton iterations — for 34 random-action env-steps with
    25-step episodes, so the environments are spread over flight, impact and stance (a robot that never fell is at step 9
    of its second episode, just landed; the ones that fell are anywhere), copy 256 random environments'
    complete state into oracle environments, take one more env-step on both sides and compare those 256: done /
    truncation bit-exact, obs / reward / info within the stated tolerances, outliers must be confirmed decision flips."""
    from opendog_b200.env import BatchedWalkEnv
    from oracle.oracle import WalkEnv
    keys = ("terminal_obs", "gait_reward", "paws_in_ground", "patterns_matches", "x_position", "y_position",
            "distance_from_origin", "paw_contact_forces", "linear_vel_tracking_reward", "reward_ctrl")
    env = BatchedWalkEnv(n, seed=17, max_episode_steps=25, info_keys=keys)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(n)
    for t in range(34):
        env.step(torch.rand(n, 8, device="cuda", generator=g) * 2 - 1)
    rng = np.random.default_rng(n)
    # spread the episode clocks over 19..24 so that about one environment in six is truncated by the compared step
    # (random actions rarely tip the robot over within 34 steps, and terminal_obs / reset indexing need done envs)
    env.set_env_state(step=torch.from_numpy(rng.integers(19, 25, n).astype(np.int32)))
    ids = np.sort(rng.choice(n, 256, replace=False))
    tid = torch.from_numpy(ids).cuda()
    q, v = [x[tid].cpu().numpy() for x in env.get_state()]
    es = {k: x[tid].cpu().numpy() for k, x in env.get_env_state().items()}
    act = torch.rand(n, 8, device="cuda", generator=g) * 2 - 1
    obs, rew, done, info = env.step(act)
    obs = obs[tid].cpu().numpy(); rew = rew[tid].cpu().numpy(); done = done[tid].cpu().numpy()
    info = {k: x[tid].cpu().numpy() for k, x in info.items()}
    a = act[tid].cpu().numpy()
    outliers = n_done = n_contact = 0
    for k, gid in enumerate(ids):
        w = WalkEnv(seed=17, env_id=int(gid))
        w.e.max_steps = 25
        assert np.array_equal(np.float32(w.desired_velocity), es["desired_velocity"][k])
        w.qpos[:] = q[k]; w.qvel[:] = v[k]
        w.e.step = int(es["step"][k]); w.e.gait_index = int(es["gait_index"][k]); w.e.gait_matches = int(es["gait_matches"][k])
        w.last_action[:] = es["last_action"][k]; w.e.last_action_is_reset = int(es["fresh"][k])
        oo, orr, od, otr, otob, oi = w.step_autoreset(a[k])
        assert bool(done[k]) == od, f"env {gid}: done differs"
        n_done += int(od); n_contact += int(oi["paws_in_ground"].any())
        # (the reset obs of a done env depends on the episode counter, which the oracle copy does not share: compare the
        #  episode's last obs instead, which is what terminal_obs holds)
        mine, theirs = (info["terminal_obs"][k], otob) if od else (obs[k], oo)
        ok = (np.abs(mine - theirs).max() < 2e-4 and abs(rew[k] - orr) < 2e-4 and info["gait_reward"][k] == oi["gait_first_call"]
              and np.array_equal(info["paws_in_ground"][k], oi["paws_in_ground"]))
        if ok:
            assert_info_matches(info, oi, k, GPU_INFO)
        else:
            outliers += 1
            assert min(oi["min_gap"][0], oi["min_gap"][1]) < FLIP_M, f"env {gid}: outside tolerance without a decision flip {oi['min_gap']}"
    assert outliers <= 3, f"{outliers}/256 decision flips"
    assert n_done >= 16 and n_contact >= 64, (n_done, n_contact)


@pytest.mark.parametrize("model", ["our_robot", "go1"])
def test_fat_and_lean_step_kernels_are_bit_identical(model):
    """OdgEnvConfig::launch_fat only moves per-contact data between registers / recomputation and local memory (odg_core.cuh:
    substep<NJL, FAT>): same arithmetic, so observations, rewards, flags and states agree bit for bit, step after step,
    through contacts, terminations and auto-resets."""
    from opendog_b200.env import BatchedWalkEnv
    n = 512
    envs = [BatchedWalkEnv(n, model=model, seed=5, info_keys=None, launch_fat=f, max_episode_steps=15) for f in (0, 1)]
    o = [e.reset().clone() for e in envs]
    assert torch.equal(o[0], o[1])
    g = torch.Generator(device="cuda").manual_seed(7)
    for t in range(25):
        a = (torch.rand(n, envs[0].act_dim, device="cuda", generator=g) * 2 - 1) * (1.0 if model == "our_robot" else 0.6)
        out = [e.step(a) for e in envs]
        for k in range(3):
            assert torch.equal(out[0][k], out[1][k]), (model, t, k)
    for a_, b_ in zip(envs[0].get_state(), envs[1].get_state()):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("zero_copy", [0, 1, 2])
def test_step_host_equals_step_with_device_tensors(zero_copy):
    """`step_host` (include/odg.h: odg_step_host — host actions in, host results out, one C call) against `step` on
    CUDA tensors: same seeds, same actions, identical observations, rewards, flags and info through terminations and
    auto-resets; pageable, page-locked and non-contiguous / float64 action tensors all take the same path. With copies
    (0), with the kernel reading the page-locked actions in place (1) and writing the results in place too (2)."""
    from opendog_b200.env import BatchedWalkEnv
    n = 300
    keys = ("x_position", "paw_contact_forces", "terminal_obs")
    a_env = BatchedWalkEnv(n, seed=11, info_keys=keys, max_episode_steps=9, host_zero_copy=zero_copy)
    assert BatchedWalkEnv(n, info_keys=None).host_zero_copy == 2                 # the default
    b_env = BatchedWalkEnv(n, seed=11, info_keys=keys, max_episode_steps=9)
    assert torch.equal(a_env.reset(), b_env.reset())
    g = torch.Generator().manual_seed(3)
    for t in range(24):
        act = torch.rand(n, 8, generator=g) * 2 - 1
        host_in = act if t % 3 == 0 else (act.pin_memory() if t % 3 == 1 else act.double())      # pageable / pinned / needs conversion
        l0 = a_env.launch_count
        out = a_env.step_host(host_in, with_info=(t % 2 == 0))
        assert a_env.launch_count - l0 == 1                                      # one kernel per host step
        obs, rew, done, info = b_env.step(act.cuda())
        assert torch.equal(out[0], obs.cpu()) and torch.equal(out[1], rew.cpu())
        assert torch.equal((out[2] | out[3]).bool(), done.cpu().bool())
        if t % 2 == 0:
            for k in keys:
                assert torch.equal(out[4][k], info[k].cpu()), k
    for x, y in zip(a_env.get_state(), b_env.get_state()):
        assert torch.equal(x, y)
