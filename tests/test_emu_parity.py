"""The CUDA step kernel's SOURCE (opendog_b200/csrc/odg_core.cuh) executed by the test-only lane emulator
against the oracle — the same checks tests/test_gpu_parity.py makes on the B200, sized for the CPU suite.
Tolerances are the stated fp32 ones (see test_gpu_parity.py)."""
import numpy as np

from emu import EmuEnv
from oracle.oracle import Sim, WalkEnv


def test_single_step_parity_through_flight_impact_and_stance():
    env = EmuEnv(1, frame_skip=1, scale_actions=0, auto_reset=0)
    sim = Sim()
    sim.reset_keyframe()
    rng = np.random.default_rng(1)
    worst_q = worst_v = 0.0
    for k in range(260):
        if k % 10 == 0:
            ctrl = (np.array([2.36, -1.8] * 4) + rng.uniform(0, 1, 8) * np.array([.44, .6] * 4)).astype(np.float32)[None]
            sim.ctrl[:] = ctrl[0]
        env.set_state(sim.qpos[None], sim.qvel[None], sim.qacc_warmstart[None])
        _, _, _, _, info = env.step(ctrl)
        sim.step()
        qp, qv, _ = env.get_state()
        eq = np.abs(qp[0] - sim.qpos) - (1e-6 + 1e-5 * np.abs(sim.qpos))
        ev = np.abs(qv[0] - sim.qvel) - (1e-4 + 1e-3 * np.abs(sim.qvel))
        worst_q, worst_v = max(worst_q, eq.max()), max(worst_v, ev.max())
        assert info["ncon"][0] == sim.ncon
        fn = sum(c["force"][0] for c in sim.contacts())
        assert abs(info["fn_sum"][0] - fn) <= 1e-2 * max(1.0, fn)
    assert worst_q <= 0 and worst_v <= 0


def test_frictionless_rows_go_through_the_cone_block():
    """A model variant whose contacts are all condim 1 (the robot geoms' own setting, our_robot.xml:9, without the
    floor's condim 3): the kernel has no separate path for such rows — they are the cone block with fri = D_t = 0 —
    and must still match the oracle's scalar rows."""
    import copy
    from opendog_b200.model.compile import load_compiled
    desc = copy.deepcopy(load_compiled("our_robot"))
    for g in desc["geoms"]:
        g["condim"] = 1
    env = EmuEnv(1, model=desc, frame_skip=1, scale_actions=0, auto_reset=0)
    sim = Sim(desc)
    sim.reset_keyframe()
    rng = np.random.default_rng(4)
    worst_q = worst_v = 0.0
    for k in range(160):
        if k % 10 == 0:
            ctrl = (np.array([2.36, -1.8] * 4) + rng.uniform(0, 1, 8) * np.array([.44, .6] * 4)).astype(np.float32)[None]
            sim.ctrl[:] = ctrl[0]
        env.set_state(sim.qpos[None], sim.qvel[None], sim.qacc_warmstart[None])
        _, _, _, _, info = env.step(ctrl)
        sim.step()
        qp, qv, _ = env.get_state()
        worst_q = max(worst_q, (np.abs(qp[0] - sim.qpos) - (1e-6 + 1e-5 * np.abs(sim.qpos))).max())
        worst_v = max(worst_v, (np.abs(qv[0] - sim.qvel) - (1e-4 + 1e-3 * np.abs(sim.qvel))).max())
        assert info["ncon"][0] == sim.ncon
    assert sim.ncon >= 2 and worst_q <= 0 and worst_v <= 0, (sim.ncon, worst_q, worst_v)


def test_walk_env_and_reset_indexing():
    N = 3
    env = EmuEnv(N, seed=7, max_episode_steps=9)
    ws = [WalkEnv(seed=7, env_id=i) for i in range(N)]
    for w in ws:
        w.e.max_steps = 9
    obs = env.reset()
    oobs = np.stack([w.reset() for w in ws])
    qp, qv, _ = env.get_state()
    assert np.array_equal(qp, np.stack([w.qpos for w in ws]).astype(np.float32))     # Philox reset bit-exact
    assert np.abs(obs - oobs).max() < 1e-6
    rng = np.random.default_rng(3)
    ndone = 0
    for t in range(20):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        obs, r, term, trunc, info = env.step(a)
        res = [w.step_autoreset(a[i]) for i, w in enumerate(ws)]
        done = term | trunc
        assert np.array_equal(done, np.array([x[2] for x in res]))
        ndone += done.sum()
        assert np.abs(obs - np.stack([x[0] for x in res])).max() < 1e-4
        assert np.abs(r - np.array([x[1] for x in res])).max() < 1e-4
        assert np.array_equal(info["gait_reward"], [x[5]["gait_first_call"] for x in res])
        assert np.array_equal(info["paws_in_ground"], np.stack([x[5]["paws_in_ground"] for x in res]))
        for i in range(N):
            if done[i]:
                assert np.abs(info["terminal_obs"][i] - res[i][4]).max() < 1e-4
        qp, qv, _ = env.get_state()
        for i, w in enumerate(ws):                      # re-sync: per-step agreement, not chaos
            w.qpos[:] = qp[i]; w.qvel[:] = qv[i]
    assert ndone == 2 * N


# An env-step may differ from the oracle beyond the stated tolerance ONLY when the fp32 kernel and the fp64 oracle took a
# different DISCRETE decision in one of its substeps, and the oracle itself reports that decision as a coin toss at fp32
# resolution: a support point within `FLIP_M` of the contact margin (the contact exists on one side only), or a support
# vertex leading its runner-up by less than `FLIP_M` (the contact sits on a different hull vertex). Coordinates are
# ~0.1 m, one float32 ulp there is 7.5e-9 m and the kinematic chain accumulates a few dozen roundings.
FLIP_M = 1e-6


def assert_info_matches(info, oi, i, names=None):
    """`info` outputs of env i (reward_calc.py:339-370, WalkEnvironment.py:65-72) against the oracle's, stated tolerances."""
    n = names or dict(paw="paw_forces", lin="lin_vel_reward", dist="distance")
    assert abs(info["x_position"][i] - oi["x_position"]) <= 1e-5
    assert abs(info["y_position"][i] - oi["y_position"]) <= 1e-5
    assert abs(info[n["dist"]][i] - oi["distance_from_origin"]) <= 1e-5
    assert abs(info[n["lin"]][i] - oi["linear_vel_tracking_reward"]) <= 2e-4
    assert abs(info["reward_ctrl"][i] - oi["reward_ctrl"]) <= 1e-3 * max(1.0, oi["reward_ctrl"])
    assert float(info["patterns_matches"][i]) == oi["patterns_matches"]
    # get_paw_contact_forces (quirk C9): the force of the LAST contact of each paw in the calf frame. The split of a paw's
    # load over its 3-4 coplanar contact points is statically indeterminate (only the regulariser picks it), so single
    # contact forces carry a looser tolerance than their sum (1 %, test_single_step_*): 5 % of the largest paw force.
    pf = np.asarray(info[n["paw"]][i], dtype=np.float64).reshape(4, 6)
    ref = oi["paw_contact_forces"]
    assert np.abs(pf - ref).max() <= 5e-2 * max(1.0, np.abs(ref).max()), (pf, ref)
    assert not pf[:, 3:].any() and not ref[:, 3:].any()              # torque slots stay zero (mj_contactForce, condim 3)
    for k in range(4):                                               # a paw off the ground reports zeros
        if not oi["paws_in_ground"][k]:
            assert not pf[k].any()


def test_walk_info_outputs_and_outliers_are_decision_flips():
    """Rows a9 / a4 of SURVEY section 8: every `info` output against the oracle, and every env-step outside the stated
    tolerance must be explained by a decision the oracle reports as within FLIP_M of flipping."""
    N, T = 6, 32
    env = EmuEnv(N, seed=11, max_episode_steps=25)                   # (the reset drops the robot: it lands at step ~8)
    ws = [WalkEnv(seed=11, env_id=i) for i in range(N)]
    for w in ws:
        w.e.max_steps = 25
    env.reset(); [w.reset() for w in ws]
    rng = np.random.default_rng(5)
    n_out = n_contact_steps = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        obs, r, term, trunc, info = env.step(a)
        res = [w.step_autoreset(a[i]) for i, w in enumerate(ws)]
        for i in range(N):
            oobs, orew, odone, _, otob, oi = res[i]
            assert (term[i] or trunc[i]) == odone
            mine = info["terminal_obs"][i] if odone else obs[i]
            theirs = otob if odone else oobs
            ok = (np.abs(mine - theirs).max() < 2e-4 and abs(r[i] - orew) < 2e-4
                  and np.array_equal(info["paws_in_ground"][i], oi["paws_in_ground"]))
            if ok:
                assert_info_matches(info, oi, i)
                n_contact_steps += int(oi["paws_in_ground"].any())
            else:
                n_out += 1
                assert min(oi["min_gap"][0], oi["min_gap"][1]) < FLIP_M, (t, i, oi["min_gap"])
        qp, qv, _ = env.get_state()
        for i, w in enumerate(ws):
            w.qpos[:] = qp[i]; w.qvel[:] = qv[i]
    assert n_out <= 2 and n_contact_steps > N * T // 3


def test_results_do_not_depend_on_sharding():
    """Envs are keyed by GLOBAL id: one handle of 4 envs == two handles of 2 (rank 0 / rank 1 shards)."""
    whole = EmuEnv(4, seed=9)
    parts = [EmuEnv(2, seed=9, first_env_id=0), EmuEnv(2, seed=9, first_env_id=2)]
    o = whole.reset()
    po = np.concatenate([p.reset() for p in parts])
    assert np.array_equal(o, po)
    rng = np.random.default_rng(0)
    for t in range(3):
        a = rng.uniform(-1, 1, (4, 8)).astype(np.float32)
        o, r, te, tr, _ = whole.step(a, info=False)
        outs = [p.step(a[2 * k:2 * k + 2], info=False) for k, p in enumerate(parts)]
        assert np.array_equal(o, np.concatenate([x[0] for x in outs]))
        assert np.array_equal(r, np.concatenate([x[1] for x in outs]))


def test_go1_single_step_parity():
    """The second model descriptor (Unitree Go1: 3 joints per leg, joint damping -> implicit-damping Euler, sphere
    feet; go1.xml) through the same kernel source, against the oracle: stand-up from the keyframe, then random targets."""
    env = EmuEnv(1, model="go1", frame_skip=1, scale_actions=0, auto_reset=0, solver_iterations=100)   # MuJoCo's cap
    sim = Sim("go1")
    sim.reset_keyframe()
    rng = np.random.default_rng(2)
    lo = np.array([r[0] for r in sim.desc["act_ctrlrange"]]); hi = np.array([r[1] for r in sim.desc["act_ctrlrange"]])
    ctrl = np.array(sim.desc["key_ctrl"], dtype=np.float32)[None]
    worst_q = worst_v = 0.0
    for k in range(220):
        if k >= 60 and k % 40 == 0:
            mid = np.array(sim.desc["key_ctrl"])
            ctrl = np.clip(mid + rng.uniform(-0.15, 0.15, 12), lo, hi).astype(np.float32)[None]
        sim.ctrl[:] = ctrl[0]
        env.set_state(sim.qpos[None], sim.qvel[None], sim.qacc_warmstart[None])
        _, _, _, _, info = env.step(ctrl)
        sim.step()
        qp, qv, _ = env.get_state()
        eq = np.abs(qp[0] - sim.qpos) - (1e-6 + 1e-5 * np.abs(sim.qpos))
        ev = np.abs(qv[0] - sim.qvel) - (1e-4 + 1e-3 * np.abs(sim.qvel))
        worst_q, worst_v = max(worst_q, eq.max()), max(worst_v, ev.max())
        assert info["ncon"][0] == sim.ncon, k
    assert worst_q <= 0 and worst_v <= 0, (worst_q, worst_v)
    assert sim.ncon >= 3                      # it is standing on its feet


def test_go1_observation_layout_48():
    """OdgEnvConfig::obs_layout = 1, the 48-value observation of the 12-actuator model (BASELINE configs[2]):
    landing_environment.py:116-136 plus the desired velocity, with the reference's get_projected_gravity formula
    (landing_environment_reward_calc.py:88-98) restated in numpy here."""
    e48 = EmuEnv(4, model="go1", seed=3, obs_layout=1)
    e45 = EmuEnv(4, model="go1", seed=3)
    rng = np.random.default_rng(0)
    o48, o45 = e48.reset(), e45.reset()
    for t in range(4):
        assert o48.shape == (4, 48) and o45.shape == (4, 45)
        assert np.array_equal(o48[:, :6], o45[:, :6]) and np.array_equal(o48[:, 9:12], o45[:, 6:9])
        assert np.array_equal(o48[:, 24:], o45[:, 21:])                   # joint velocities, last action
        q, _, _ = e48.get_state()
        key = np.array(e48.desc["key_qpos"][7:19], np.float32)
        assert np.allclose(o48[:, 12:24], q[:, 7:19] - key, atol=1e-6)
        w, x, y, z = q[:, 3:7].astype(np.float64).T
        eul = np.stack([np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y)),
                        np.arcsin(np.clip(2 * (w * y - z * x), -1, 1)),
                        np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))], 1)
        v = (eul @ np.array(e48.desc["gravity"]))[:, None] * eul
        n = np.linalg.norm(v, axis=1, keepdims=True)
        assert np.allclose(o48[:, 6:9], np.where(n == 0, v, v / np.where(n == 0, 1, n)), atol=1e-6)
        a = rng.uniform(-0.3, 0.3, (4, 12)).astype(np.float32)
        o48, o45 = e48.step(a, info=False)[0], e45.step(a, info=False)[0]


def test_support_vertex_candidate_lists_contain_every_winner():
    """The collision pass scans only the hull vertices listed for the direction cell of the floor direction
    (odg_prep.h: hull_vertex_is_candidate; 24 cells: dominant axis x signs). For random link orientations the first-index arg-extremum over the list
    must be the one over the whole hull, for the floor direction and the three tilted ones — in float32, with the
    kernel's expressions — and the lists must actually prune."""
    import ctypes as C
    from emu import lib
    env = EmuEnv(1)
    L = lib()
    L.emu_support_candidates.restype = C.c_int
    L.emu_num_slots.restype = C.c_int
    L.emu_tilt_dir.restype = C.c_float
    h = C.c_void_p(env.h)
    tilt = np.array([[L.emu_tilt_dir(h, i, c) for c in range(3)] for i in range(3)], np.float32)
    rng = np.random.default_rng(0)
    total = listed = 0
    for slot in range(L.emu_num_slots(h)):
        for leg in range(4):
            verts = np.zeros((400, 3), np.float32); cand = np.zeros(400, np.int32); n = C.c_int()
            lists = []
            for o in range(24):
                nv = L.emu_support_candidates(h, slot, leg, o, verts.ctypes.data_as(C.c_void_p), cand.ctypes.data_as(C.c_void_p), C.byref(n))
                lists.append(cand[:n.value].copy())
                assert np.all(np.diff(lists[-1]) > 0)                      # index order, no duplicates
                total += nv; listed += n.value
            V = verts[:nv]
            # random rotations (QR of a Gaussian matrix), plus near-axis-aligned ones where octants meet
            for trial in range(300):
                A = rng.normal(size=(3, 3)) if trial % 3 else np.eye(3) + 0.02 * rng.normal(size=(3, 3))
                R, _ = np.linalg.qr(A)
                R = (R * np.sign(np.linalg.det(R))).astype(np.float32)
                rz = R[2]
                ar = np.abs(rz)
                axis = 0 if (ar[0] >= ar[1] and ar[0] >= ar[2]) else (1 if ar[1] >= ar[2] else 2)
                o = axis * 8 + (int(rz[0] > 0) | int(rz[1] > 0) << 1 | int(rz[2] > 0) << 2)
                c = lists[o]
                z = V @ rz
                assert c[np.argmin(z[c])] == np.argmin(z), (slot, leg, o)
                for i in range(3):
                    sc = V @ (R.T @ tilt[i])
                    assert c[np.argmax(sc[c])] == np.argmax(sc), (slot, leg, o, i)
    assert listed < 0.35 * total, (listed, total)


def test_go1_colliders_condim6_and_cfrc_ext_tumble_parity():
    """The full Go1 collision set (unitree_go1/go1.xml:26-64: trunk boxes / cylinders / capsules, hip cylinders, thigh and
    calf capsules, condim-6 sphere feet with torsional and rolling friction) through the kernel source against the oracle:
    robots dropped in random orientations with random targets tumble over every collider type. Single mj_step from
    identical states at the stated tolerance; data.cfrc_ext (mj_rnePostConstraint) of the same forward pass."""
    env = EmuEnv(1, model="go1", frame_skip=1, scale_actions=0, auto_reset=0, solver_iterations=100)
    sim = Sim("go1")
    rng = np.random.default_rng(5)
    lo = np.array([r[0] for r in sim.desc["act_ctrlrange"]]); hi = np.array([r[1] for r in sim.desc["act_ctrlrange"]])
    geoms = sim.desc["geoms"]
    assert len(geoms) == 42 and sum(g["condim"] == 6 for g in geoms) == 4
    seen, bad, n, worst_cf = set(), [], 0, 0.0
    for trial in range(5):
        sim.reset_keyframe()
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        sim.qpos[3:7] = q; sim.qpos[2] = 0.45
        sim.qpos[7:] = rng.uniform(lo, hi)
        sim.qvel[:] = rng.normal(size=18) * 0.5
        for k in range(260):
            if k % 50 == 0:
                ctrl = rng.uniform(lo, hi).astype(np.float32)[None]
            sim.ctrl[:] = ctrl[0]
            env.set_state(sim.qpos[None], sim.qvel[None], sim.qacc_warmstart[None])
            _, _, _, _, info = env.step(ctrl)
            sim.step()
            qp, qv, _ = env.get_state()
            eq = (np.abs(qp[0] - sim.qpos) - (1e-6 + 1e-5 * np.abs(sim.qpos))).max()
            ev = (np.abs(qv[0] - sim.qvel) - (1e-4 + 1e-3 * np.abs(sim.qvel))).max()
            n += 1
            for c in sim.contacts():
                seen.add((geoms[c["geom"]]["type"], geoms[c["geom"]]["leg"] < 0, c["dim"]))
            flip = sim.decision_gaps[0] < 1e-6            # a support point within fp32 rounding of the contact margin
            if info["ncon"][0] != sim.ncon:
                assert flip, (trial, k, info["ncon"][0], sim.ncon, sim.decision_gaps)
                continue
            if eq > 0 or ev > 0:
                bad.append((trial, k, float(eq), float(ev)))
            cf = sim.cfrc_ext()[1:]
            scale = 1.0 + np.abs(cf).max()
            worst_cf = max(worst_cf, np.abs(info["cfrc_ext"][0].reshape(13, 6) - cf).max() / scale)
    # every collider type on trunk and legs, and the 6-row feet, were exercised
    assert {(1, False, 6), (2, False, 3), (2, True, 3), (3, False, 3), (3, True, 3), (4, True, 3)} <= seen, seen
    # stiff multi-point rests (a cylinder rim's four points next to capsule ends) converge a little short of the stated
    # tolerance in fp32 now and then: at most 0.5 % of the steps, never by more than 10x
    assert len(bad) <= 0.005 * n and all(b[2] < 1e-5 and b[3] < 1e-3 for b in bad), bad
    assert worst_cf < 2e-2, worst_cf


def test_jump_task_matches_the_oracle():
    """JumpEnvironmentV0 on the Go1 model (environments/JumpEnvironment.py:70-134, rewards/jump_environment_reward_calc.py)
    through the kernel source: reset observation and Philox reset state bit-for-bit, then on IDENTICAL states (odg_evaluate
    semantics) flags bit-exact and every weighted reward term within float32 of the oracle, which is pinned to the
    reference's own code by tests/test_golden_jump.py."""
    from oracle.go1_tasks import JumpEnv
    N, seed = 4, 11
    env = EmuEnv(N, model="go1", seed=seed, task=1, scale_actions=0, reset_noise_scale=0.1, auto_reset=0, solver_iterations=100)
    ws = [JumpEnv(seed=seed, env_id=i) for i in range(N)]
    obs = env.reset()
    oobs = np.stack([w.reset() for w in ws])
    assert obs.shape == (N, 21) and np.abs(obs - oobs).max() < 1e-6
    assert np.array_equal(env.get_state()[0], np.stack([w.qpos for w in ws]).astype(np.float32))
    assert np.array_equal(env.get_env_state()["desired_velocity"], np.stack([w.desired_velocity for w in ws]))
    rng = np.random.default_rng(3)
    lo = np.array([r[0] for r in ws[0].sim.desc["act_ctrlrange"]]); hi = np.array([r[1] for r in ws[0].sim.desc["act_ctrlrange"]])
    key = np.array(ws[0].sim.desc["key_ctrl"])
    n_cc = n_air = n_term = 0
    for t in range(40):
        ctrl = np.clip(key + rng.uniform(-0.5, 0.5, (N, 12)), lo, hi).astype(np.float32)
        for i, w in enumerate(ws):                       # walk the oracle, with poses that exercise every term
            w.step(ctrl[i])
            if i == 1 and t % 5 == 2:
                w.qpos[:3] = [rng.uniform(0.2, 1.4), rng.uniform(-0.4, 0.4), rng.uniform(0.48, 0.7)]
            if i == 2 and t % 6 == 3:
                w.qpos[2] = 0.13; w.qpos[3:7] = [np.cos(0.725), np.sin(0.725), 0.0, 0.0]
            if i == 3 and t % 9 == 4:
                a = 0.36 if t % 2 else -0.34                # around the +-20 degree yaw bound
                w.qpos[3:7] = [np.cos(a / 2), 0.0, 0.0, np.sin(a / 2)]
            w.qpos[:] = w.qpos.astype(np.float32); w.qvel[:] = w.qvel.astype(np.float32)   # float32-representable states
        env.set_state(np.stack([w.qpos for w in ws]), np.stack([w.qvel for w in ws]),
                      np.stack([w.sim.qacc_warmstart for w in ws]))
        env.set_env_state(step=np.array([w.step_count for w in ws], np.int32))
        obs, rew, term, trunc, info = env.evaluate(ctrl)
        for i, w in enumerate(ws):
            w.sim.ctrl[:] = ctrl[i]
            oo, orew, oterm, otrunc, oinfo = w.evaluate()
            assert np.abs(obs[i] - oo).max() < 2e-6, (t, i)
            near = abs(oinfo["collision_norm"] - 0.1) < 1e-3
            assert term[i] == oterm and trunc[i] == otrunc, (t, i)
            if not near:
                assert np.allclose(info["task_terms"][i], oinfo["terms"], rtol=3e-6, atol=1e-6), (t, i, info["task_terms"][i], oinfo["terms"])
                assert abs(rew[i] - orew) <= 3e-6 * max(1.0, abs(orew)), (t, i)
                assert abs(info["reward_raw"][i] - oinfo["reward_unclipped"]) <= 1e-5 * max(1.0, abs(oinfo["reward_unclipped"]))
            n_cc += oinfo["terms"][9] > 0; n_air += oinfo["terms"][0] > 0; n_term += oterm
    assert n_cc >= 4 and n_air >= 4 and n_term >= 4
