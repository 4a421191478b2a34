"""The SB3-facing adapter `opendog_b200.compat.WalkVecEnv` (SURVEY section 8 row a12 / boundary surface S2): what
`train/train.py:63-87,117-158` hands to `PPO(...)`. CPU part: the lazy info mapping and that the class really IS a
`stable_baselines3.common.vec_env.VecEnv` when SB3 is importable (checked with a stand-in module, SB3 is not in this
image). GPU part: the call sequence of SB3's `collect_rollouts` against the plain batched env."""
import importlib
import sys
import types

import numpy as np
import pytest


def test_lazy_info_behaves_like_the_dict_sb3_expects():
    from opendog_b200.compat import LazyInfo
    rows = {"x_position": np.arange(4, dtype=np.float32), "patterns_matches": np.full(4, 16, np.float32),
            "paw_contact_forces": np.arange(4 * 4 * 6, dtype=np.float32).reshape(4, 4, 6)}
    d = LazyInfo(rows, 2)
    assert "x_position" in d and d["x_position"] == 2.0 and isinstance(d["x_position"], float)
    assert d.get("missing", 7) == 7 and "missing" not in d
    with pytest.raises(KeyError):
        d["missing"]
    pf = d["paw_contact_forces"]                       # {paw body id: float64[6]} as reward_calc.py:351-370 returns it
    assert sorted(pf) == [4, 7, 10, 13] and pf[7].dtype == np.float64 and np.array_equal(pf[7], rows["paw_contact_forces"][2, 1])
    d["episode"] = {"r": 1.0}                          # VecMonitor-style assignment
    assert "episode" in d and set(d.keys()) == {"x_position", "patterns_matches", "paw_contact_forces", "episode"}
    c = d.copy()
    assert type(c) is dict and c["x_position"] == 2.0 and c["episode"] == {"r": 1.0}
    # the reference's CustomLoggingCallback loop (train/train.py:33-37)
    got = [d[k] for k in ("x_position", "y_position", "patterns_matches") if k in d]
    assert got == [2.0, 16.0]


def test_walkvecenv_subclasses_sb3_vecenv_when_sb3_is_importable(monkeypatch):
    """SB3's `_wrap_env` only accepts instances of its own VecEnv; a duck-typed class is wrapped as a single gym env and
    rejected. With an importable `stable_baselines3.common.vec_env.VecEnv` the adapter must derive from it."""
    class FakeVecEnv:
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space
            self.sb3_init_called = True
    sb3 = types.ModuleType("stable_baselines3"); common = types.ModuleType("stable_baselines3.common")
    vec = types.ModuleType("stable_baselines3.common.vec_env"); vec.VecEnv = FakeVecEnv
    sb3.common = common; common.vec_env = vec
    for name, mod in (("stable_baselines3", sb3), ("stable_baselines3.common", common), ("stable_baselines3.common.vec_env", vec)):
        monkeypatch.setitem(sys.modules, name, mod)
    import opendog_b200.compat as compat
    try:
        compat = importlib.reload(compat)
        assert issubclass(compat.WalkVecEnv, FakeVecEnv)
        for m in ("reset", "step_async", "step_wait", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped",
                  "get_images", "render"):
            assert callable(getattr(compat.WalkVecEnv, m)), m
    finally:
        for name in ("stable_baselines3", "stable_baselines3.common", "stable_baselines3.common.vec_env"):
            monkeypatch.delitem(sys.modules, name, raising=False)
        importlib.reload(compat)
    assert not issubclass(compat.WalkVecEnv, FakeVecEnv)


@pytest.mark.gpu
def test_collect_rollouts_call_sequence_against_the_batched_env():
    torch = pytest.importorskip("torch")
    from opendog_b200.compat import WalkVecEnv
    from opendog_b200.env import BatchedWalkEnv
    N, T = 48, 40
    kw = dict(seed=6, max_episode_steps=12)
    venv = WalkVecEnv(N, **kw)
    ref = BatchedWalkEnv(N, auto_reset=1, info_keys=("terminal_obs", "x_position", "patterns_matches", "paw_contact_forces"), **kw)
    assert venv.num_envs == N and venv.observation_space.shape == (33,) and venv.action_space.shape == (8,)
    assert venv.observation_space.dtype == np.float64 and venv.action_space.dtype == np.float32
    assert venv.get_attr("render_mode") == ["rgb_array"] * N and venv.env_is_wrapped(object) == [False] * N
    assert venv.get_attr("metadata", indices=[0])[0]["render_fps"] == 50
    obs = venv.reset()
    robs = ref.reset()
    assert obs.dtype == np.float64 and obs.shape == (N, 33) and np.array_equal(obs, robs.double().cpu().numpy())
    rng = np.random.default_rng(0)
    ret = np.zeros(N); length = np.zeros(N, int)
    n_eps = 0
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)           # SB3 clips the policy output to the action space
        venv.step_async(a)
        obs, rew, done, infos = venv.step_wait()
        ro, rr, rd, rinfo = ref.step(torch.from_numpy(a).cuda())
        assert obs.dtype == np.float64 and rew.dtype == np.float64 and done.dtype == bool and len(infos) == N
        assert np.array_equal(obs, ro.double().cpu().numpy()) and np.array_equal(rew, rr.double().cpu().numpy())
        assert np.array_equal(done, rd.cpu().numpy())
        ret += rew; length += 1
        tobs = rinfo["terminal_obs"].cpu().numpy(); rx = rinfo["x_position"].cpu().numpy()
        trunc = ref.truncated.cpu().numpy().astype(bool); term = ref.terminated.cpu().numpy().astype(bool)
        for i, info in enumerate(infos):
            assert info["x_position"] == float(rx[i]) and "distance_from_origin" in info and "y_position" in info
            if done[i]:
                n_eps += 1
                assert np.array_equal(info["terminal_observation"], tobs[i].astype(np.float64))
                assert info["TimeLimit.truncated"] == bool(trunc[i] and not term[i])
                ep = info["episode"]                                   # Monitor(info_keywords=...) of train/train.py:69-70,101
                assert ep["l"] == length[i] and abs(ep["r"] - ret[i]) < 1e-5 and ep["t"] >= 0
                assert ep["x_position"] == info["x_position"] and sorted(ep["paw_contact_forces"]) == [4, 7, 10, 13]
                assert ep["patterns_matches"] == info["patterns_matches"]
                ret[i] = 0; length[i] = 0
            else:
                assert "terminal_observation" not in info and "episode" not in info
    assert n_eps >= 3 * N
    # VecEnv.step = step_async + step_wait; env_method / set_attr refuse what a batched simulator cannot do
    out = venv.step(np.zeros((N, 8), np.float32))
    assert len(out) == 4
    with pytest.raises(NotImplementedError):
        venv.env_method("some_method")
    venv.set_attr("render_mode", None); assert venv.get_attr("render_mode")[0] is None
    assert venv.get_images() == [None] * N
    venv.close(); ref.close()
