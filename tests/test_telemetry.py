"""Telemetry packet schema (wireless_comunication/server.py:95-118) through a real UDP socket on localhost."""
import socket

import pytest

torch = pytest.importorskip("torch")


@pytest.mark.gpu
def test_udp_packet_matches_reference_schema():
    import msgpack
    from opendog_b200.env import BatchedWalkEnv
    from opendog_b200.telemetry import TelemetryServer
    env = BatchedWalkEnv(8, seed=1, info_keys=("paw_contact_forces", "ncon"))
    env.reset()
    for _ in range(12):
        env.step(torch.zeros(8, 8, device="cuda"))
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
    srv.close(); cli.close()
