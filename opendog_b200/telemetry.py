"""UDP + MessagePack telemetry of one environment of a batched simulator: the packet
`wireless_comunication/server.py:95-118` sends and `client.py` plots, so the reference's live paw-force plotter keeps
working against the GPU simulator. Same dict keys, same value layouts:
    timestamp, num_qpos, num_qvel, num_act, qpos_data[3], qvel_data[3], ctr_data[nu],
    contact_forces_data[4*6] (paw order 4, 7, 10, 13 = FL FR BL BR), active_contacts
"""
from __future__ import annotations

import socket
import time

import msgpack


def packet(env, index: int = 0, ctrl=None) -> dict:
    """The reference's `_get_simulation_data` dict for environment `index` of a BatchedWalkEnv whose info_keys include
    `paw_contact_forces` (and `ncon` for active_contacts)."""
    qpos, qvel = env.get_state()
    info = env.info
    forces = info["paw_contact_forces"][index].reshape(-1).double().cpu().tolist() if "paw_contact_forces" in info else [0.0] * 24
    ncon = int(info["ncon"][index]) if "ncon" in info else 0
    if ctrl is None:
        ctrl = env.get_env_state()["last_action"][index].double().cpu().tolist()     # last commanded targets (rad)
    return {
        "timestamp": time.time(), "num_qpos": env.nq, "num_qvel": env.nv,
        "num_act": 0,                                   # `m.na` in the reference: activation states, none in this model
        "qpos_data": qpos[index, :3].double().cpu().tolist(), "qvel_data": qvel[index, :3].double().cpu().tolist(),
        "ctr_data": list(ctrl), "contact_forces_data": forces, "active_contacts": ncon,
    }


class TelemetryServer:
    """Minimal stand-in for `MujocoCommunicationServer` (server.py:19-66): remembers the last client that sent a
    datagram to (host, port) and answers `send(env)` with one msgpack datagram."""

    def __init__(self, host: str = "127.0.0.1", port: int = 12345):
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.sock.bind((host, port))
        self.sock.settimeout(0.0)
        self.client = None
        self.address = self.sock.getsockname()

    def poll(self):
        try:
            _, addr = self.sock.recvfrom(1024)
            self.client = addr
        except (BlockingIOError, socket.timeout):
            pass
        return self.client

    def send(self, env, index: int = 0):
        self.poll()
        if self.client is None:
            return False
        self.sock.sendto(msgpack.packb(packet(env, index)), self.client)
        return True

    def close(self):
        self.sock.close()


class CustomMetrics:
    """The scalars `train/train.py:20-44` (`CustomLoggingCallback`) writes to TensorBoard — the mean of `x_position`,
    `y_position`, `distance_from_origin` and `patterns_matches` over every environment and every step since the last
    record, recorded every 100 calls — for a batched environment: the per-step `info` tensors are summed on the device
    (one small kernel per key and step, no host copy), and `on_step` returns `{"custom/<key>": mean}` once every `every`
    calls (one 4-float D2H) and `None` otherwise. `record` is called with that dict if given (e.g. an SB3 logger's
    `record` wrapped as `lambda d: [logger.record(k, v) for k, v in d.items()]`)."""

    KEYS = ("x_position", "y_position", "distance_from_origin", "patterns_matches")

    def __init__(self, every: int = 100, keys=KEYS, record=None):
        self.every, self.keys, self.record = int(every), tuple(keys), record
        self.n_calls = 0
        self._sum = None
        self._count = 0

    def on_step(self, info: dict):
        import torch
        self.n_calls += 1
        vals = [info[k] for k in self.keys if k in info]
        if vals:
            present = [k for k in self.keys if k in info]
            if self._sum is None or self._present != present:
                self._present = present
                self._sum = torch.zeros(len(present), dtype=torch.float64, device=vals[0].device)
                self._count = 0
            self._sum += torch.stack([v.reshape(-1).double().sum() for v in vals])
            self._count += vals[0].numel()
        if self.n_calls % self.every != 0 or self._sum is None or self._count == 0:
            return None
        means = (self._sum / self._count).cpu().tolist()
        out = {f"custom/{k}": m for k, m in zip(self._present, means)}
        self._sum.zero_(); self._count = 0
        if self.record is not None:
            self.record(out)
        return out
