#!/usr/bin/env python
"""Solver statistics from the host lane emulator (no GPU): per Newton iteration the relative step size, the accepted
step length, the line-search passes and the contact count of lane 0's leg — the data behind DESIGN.md's divergence
analysis. Builds its own instrumented copy of the emulator (-DODG_EMU_STATS) under /tmp.

    python tools/emu_solver_stats.py [envs] [env_steps] [extra -D flags ...]
"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu")); sys.path.insert(0, ROOT)
import emu as E

def build(flags, tag):
    so = f"/tmp/libodg_emu_stats_{tag}.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-DODG_EMU_STATS",
                           *flags, "-o", so, os.path.join(ROOT, "tests/emu/odg_emu.cpp")])
    return so

def run(N, steps, flags, tag="a", seed=0):
    so = build(flags, tag)
    L = C.CDLL(so)
    L.emu_create.restype = C.c_void_p
    L.emu_create.argtypes = [C.POINTER(E.OdgModel), C.POINTER(E.OdgEnvConfig), C.c_int, C.c_uint64]
    for name in ("emu_destroy", "emu_reset", "emu_step", "emu_get_state", "emu_set_state", "emu_get_env_state", "emu_set_env_state"):
        getattr(L, name).restype = None
    E._lib = L
    env = E.EmuEnv(N, seed=seed, frame_skip=1, max_episode_steps=7500, auto_reset=1)
    env.reset()
    rng = np.random.default_rng(seed)
    buf = np.zeros(6 * 400 * N, np.float32)
    recs, its, states = [], [], []
    for t in range(steps * 10):
        if t % 10 == 0:
            a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        obs, rew, term, trunc, info = env.step(a)
        n = L.emu_stats(buf.ctypes.data_as(C.c_void_p), buf.size)
        if t >= 120:
            recs.append(buf[:n].reshape(-1, 6).copy()); its.append(info["solver_iters"].copy())
    q, v, w = env.get_state()
    return np.concatenate(recs), np.array(its), (q, v)

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    flags = sys.argv[3:]
    r, its, _ = run(N, steps, flags)
    it, rel, al, ps, nc, d10 = r.T
    print(f"flags {flags}: iterations/substep mean {its.mean():.3f} p90 {np.percentile(its, 90):.0f} max {its.max()}  "
          f"warp-max(8) mean {its[:, :N // 8 * 8].reshape(len(its), -1, 8).max(2).mean():.3f}")
    print(f"  ls passes/iteration {ps.sum() / len(ps):.3f}; passes hist {np.bincount(ps.astype(int))}")
    acc = ps > 0
    print(f"  accepted alpha quantiles (with LS): {np.percentile(al[acc], [1, 10, 50, 90, 99])}")
    print(f"  first-pass accept: {np.mean(ps[acc] == 1):.3f}; alpha==1 {np.mean(al[acc] == 1):.3f} alpha==0.5 {np.mean(al[acc] == 0.5):.3f} "
          f"alpha==2 {np.mean(al[acc] == 2):.3f} alpha==.25 {np.mean(al[acc] == .25):.3f}")
    print(f"  rel step of iterations: log10 hist", np.histogram(np.log10(np.maximum(rel, 1e-12)), bins=[-12, -6, -5, -4, -3, -2, -1, 0, 3])[0])
    # the iteration before the last one of each substep
    last = np.r_[it[1:] == 0, True]
    prev = np.r_[last[1:], False] & ~last
    print(f"  rel step of the second-to-last iteration: quantiles {np.percentile(rel[prev], [10, 50, 90])}; of the last {np.percentile(rel[last], [10, 50, 90])}")
    print(f"  second-to-last: alpha==1 {np.mean(al[prev] == 1):.3f}, passes==1 {np.mean(ps[prev] == 1):.3f}")
