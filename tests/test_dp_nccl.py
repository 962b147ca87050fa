"""Data-parallel correctness on hardware (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl.py -m gpu`).
See tests/dp_worker.py for what is checked; reference: README.md:93-94, transfer_em/cgan.py:8-11."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("is3d", [True, False])
def test_dp_gradients_equal_global_batch(is3d):
    world = 2
    env = dict(os.environ, TEM_DP_3D="1" if is3d else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631" if is3d else "29632", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DP_RESULT ")][-1]
    res = json.loads(line[len("DP_RESULT "):])
    print(res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dp_nccl_{'3d' if is3d else '2d'}.json"), "w") as f:
        json.dump(res, f)
    assert res["replicas_bit_identical"] and res["step"] == 3
    # fp32 atomics order differs between a batch-1 and a batch-2 launch: summation noise, not a scaling error
    # (measured on 2 x B200, profiles/dp_nccl_r2.log: 3-D 3e-7, 2-D 3e-5 on the direct-kernel discriminators)
    assert max(res["grad_rel_l2"].values()) < (1e-5 if is3d else 1e-4), res
    assert res["loss_rel"] < 1e-5, res
    assert max(res["params_vs_single_rel_l2"].values()) < 2e-2, res      # Adam's sign-like first steps amplify 1e-5 gradient noise
