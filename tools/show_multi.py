#!/usr/bin/env python
"""Print the headline and configs[3] numbers of a bench.py JSON line (multi-GPU runs)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus %d value %.4e e2e %.4e" % (d["n_gpus"], d["value"], d["e2e"]["value"]))
c = d.get("configs3") or {}
print({k: c.get(k) for k in ("rollout_env_steps_per_s", "train_iteration_env_steps_per_s", "ms_rollout", "ms_ppo_epoch", "ms_grad_allreduce", "ms_gae_and_stats_allreduce")})
