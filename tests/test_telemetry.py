"""Telemetry packet schema (wireless_comunication/server.py:95-118) through a real UDP socket on localhost."""
import socket

import pytest

torch = pytest.importorskip("torch")


@pytest.mark.gpu
def test_udp_packet_matches_reference_schema():
    import msgpack
    from opendog_b200.env import BatchedWalkEnv
    from opendog_b200.telemetry import TelemetryServer
    import numpy as np
    from oracle.oracle import WalkEnv, scale_action
    env = BatchedWalkEnv(8, seed=1, info_keys=("paw_contact_forces", "ncon"))
    env.reset()
    for _ in range(11):
        env.step(torch.zeros(8, 8, device="cuda"))
    # the VALUES of the packet, not only its schema: an oracle copy of environment 3 takes the last step alongside
    q, v = [x.cpu().numpy() for x in env.get_state()]
    w = WalkEnv(seed=1, env_id=3); w.reset()
    w.qpos[:] = q[3]; w.qvel[:] = v[3]
    act = np.linspace(-0.6, 0.6, 8).astype(np.float32)
    env.step(torch.from_numpy(np.tile(act, (8, 1))).cuda())
    oinfo = w.step(act)[4]
    srv = TelemetryServer(port=0)
    cli = socket.socket(socket.AF_INET, socket.SOCK_DGRAM); cli.settimeout(5.0)
    cli.sendto(b"hello", srv.address)
    import time; time.sleep(0.05)
    assert srv.send(env, index=3)
    d = msgpack.unpackb(cli.recvfrom(65536)[0])
    assert set(d) == {"timestamp", "num_qpos", "num_qvel", "num_act", "qpos_data", "qvel_data", "ctr_data",
                      "contact_forces_data", "active_contacts"}
    assert (d["num_qpos"], d["num_qvel"]) == (15, 14) and len(d["qpos_data"]) == 3 and len(d["qvel_data"]) == 3
    assert len(d["ctr_data"]) == 8 and len(d["contact_forces_data"]) == 24 and d["active_contacts"] >= 4
    assert 0.03 < d["qpos_data"][2] < 0.25                     # the robot has landed and stands
    assert np.allclose(d["qpos_data"], w.qpos[:3], atol=1e-5) and np.allclose(d["qvel_data"], w.qvel[:3], atol=2e-4)
    assert np.allclose(d["ctr_data"], scale_action(act), atol=1e-6)          # data.ctrl = the scaled targets (rad)
    assert d["active_contacts"] == w.e.d.ncon
    f = np.array(d["contact_forces_data"]).reshape(4, 6); ref = oinfo["paw_contact_forces"]      # reward_calc.py:351-370
    assert np.abs(f - ref).max() <= 5e-2 * max(1.0, np.abs(ref).max()) and np.abs(ref).max() > 0.5
    srv.close(); cli.close()


def test_custom_metrics_equal_the_reference_callback_arithmetic():
    """train/train.py:20-44: every `info[key]` of every environment is appended to a list, and every 100 calls the mean
    of the list is recorded and the list cleared. Same numbers from per-step [N] tensors (CPU tensors here: the class
    only needs torch)."""
    import numpy as np
    from opendog_b200.telemetry import CustomMetrics
    rng = np.random.default_rng(0)
    N, every = 7, 5
    cm = CustomMetrics(every=every)
    lists = {k: [] for k in CustomMetrics.KEYS}
    for call in range(1, 13):
        info = {k: torch.from_numpy(rng.normal(size=N).astype(np.float32)) for k in CustomMetrics.KEYS}
        info["unrelated"] = torch.zeros(N)
        for k in CustomMetrics.KEYS:                       # the reference: one append per environment and key
            lists[k] += [float(x) for x in info[k]]
        out = cm.on_step(info)
        if call % every == 0:
            assert set(out) == {f"custom/{k}" for k in CustomMetrics.KEYS}
            for k in CustomMetrics.KEYS:
                assert abs(out[f"custom/{k}"] - np.mean(lists[k])) < 1e-6
                lists[k].clear()
        else:
            assert out is None
