"""Worker of tests/test_dp_nccl.py (one process per GPU, launched by torch.distributed.run): data-parallel correctness of
the train step (README.md:93-94, cgan.py:8-11: "normalizing loss based on global batch size").

Every rank owns one sample of a global batch of `world` samples and runs tem_train_step through tem_comm_init's NCCL
communicator; rank 0 also runs the whole batch in a second, single-process handle (tem_train_grads, no communicator).
Checked: (1) rank-averaged gradients == global-batch gradients, (2) the reported losses == global-batch losses, (3) after
3 steps every replica holds bit-identical parameters and Adam moments, which equal the single-process trajectory."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tem_oracle as O                        # noqa: E402  (test infrastructure: parameter recipe only)
from transfer_em_b200 import EM2EM                        # noqa: E402
from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY   # noqa: E402

NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    is3d = os.environ.get("TEM_DP_3D", "1") == "1"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = np.random.default_rng(61)
    P = {}
    for k in NETS:
        layers = O.generator_layers(8) if k in ('g', 'f') else O.discriminator_layers(8, is3d)
        P[k] = [p * 2.0 for p in O.init_params(layers, is3d, r)]
    shape = (world,) + (74,) * (3 if is3d else 2) + (1,)
    rx = np.clip(r.standard_normal(shape) * 0.4, -0.9, 0.9).astype(np.float32)
    ry = np.clip(r.standard_normal(shape) * 0.35 + 0.1, -0.9, 0.9).astype(np.float32)
    dp = EM2EM(74, "dp", is3d=is3d, max_batch=1, dropout=False, device=local, distributed=True, checkpoint_dir="/tmp/tem_dp_none")
    if rank == 0:                       # only rank 0 loads the recipe: the replicas must receive it through the communicator
        for k, n in NETS.items():
            dp.engine.set_weights(n, P[k])
    dp.engine._lib.tem_comm_sync_params(dp.engine._h, None); torch.cuda.synchronize()
    res = {"rank": rank, "world": world}
    losses_dp = dp.train_step(rx[rank:rank + 1], ry[rank:rank + 1])          # all-reduce + Adam inside
    g_dp = {k: dp.engine.get_vector(n, 1) / world for k, n in NETS.items()}  # arena holds the rank SUM after the all-reduce
    single = None
    if rank == 0:
        single = EM2EM(74, "single", is3d=is3d, max_batch=world, dropout=False, device=local, checkpoint_dir="/tmp/tem_dp_none")
        for k, n in NETS.items():
            single.engine.set_weights(n, P[k])
        losses_1 = single.engine.train_grads(rx, ry)
        res["grad_rel_l2"] = {k: rel_l2(g_dp[k], single.engine.get_vector(n, 1)) for k, n in NETS.items()}
        res["loss_rel"] = float(np.max(np.abs(np.array(losses_dp) - np.array(losses_1)) / np.maximum(np.abs(np.array(losses_1)), 1e-6)))
        single.engine.apply_adam(1.0)
    for _ in range(2):
        dp.train_step(rx[rank:rank + 1], ry[rank:rank + 1])
        if rank == 0:
            single.train_step(rx, ry)
    # replicas bit-identical: all-gather a checksum of params, m, v
    state = torch.cat([torch.from_numpy(dp.engine.get_vector(n, w)) for n in NETS.values() for w in (0, 2, 3)]).cuda()
    gathered = [torch.empty_like(state) for _ in range(world)]
    dist.all_gather(gathered, state)
    res["replicas_bit_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
    res["step"] = dp.engine.step
    if rank == 0:
        res["params_vs_single_rel_l2"] = {k: rel_l2(dp.engine.get_vector(n, 0) - np.concatenate([p.ravel() for p in P[k]]),
                                                     single.engine.get_vector(n, 0) - np.concatenate([p.ravel() for p in P[k]]))
                                          for k, n in NETS.items()}
        print("DP_RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
