"""CycleGAN trainer mirroring transfer_em/cgan.py (EM2EM) on top of libtem_b200.

Same constructor arguments, attributes (`generator_g/f`, `discriminator_x/y`, `buffer`, `outdimsize`,
`is3d`) and methods (`train_step`, `train`, `predict`, `make_checkpoint`) as the reference.  The four
networks, their Adam state and the whole train step live in one tem_handle; Python only passes
pointers.  Extras (keyword-only): batch size of the workspace, device, seed, loss mode, data-parallel.
"""
import glob
import json
import os
import time

import numpy as np
import torch

from ._lib import NET_G, NET_F, NET_DX, NET_DY
from .engine import Engine
from .models.utils import NetModel


class EM2EM(object):
    """Creates CGAN model for 1-channel 2d or 3d data and provides functions to train and predict.

    Compatible tensor dimension sizes: 74 (reference) and, as a superset, any n = 2 (mod 4) >= 74.
    """

    def __init__(self, dimsize, exp_name, is3d=True, norm_type="instancenorm", ckpt_restore=None, wf=8, focal_gamma=2,
                 disc_prior=None, *, max_batch=8, device=None, seed=0, dropout=True, loss_mode="focal",
                 meanstd_x=None, meanstd_y=None, distributed=False, checkpoint_dir=None, train=True):
        if dimsize < 74:
            raise RuntimeError("minimum dimension allowed is 74")                     # cgan.py:52-53
        if dimsize % 4 != 2:
            raise RuntimeError(f"{dimsize} does not allow for valid convolutions")    # generator.py:37-38
        if disc_prior is not None:
            raise NotImplementedError("disc_prior needs a Keras h5 prior model (cgan.py:21-30): out of scope")
        self.engine = Engine(dimsize=dimsize, is3d=is3d, wf=wf, max_batch=max_batch, train=train, device=device, seed=seed,
                             dropout=dropout, loss_mode=loss_mode, focal_gamma=float(focal_gamma))
        self.discriminator_x = NetModel(self.engine, NET_DX, 'discriminator')
        self.discriminator_y = NetModel(self.engine, NET_DY, 'discriminator')
        self.generator_g = NetModel(self.engine, NET_G, 'generator')
        self.generator_f = NetModel(self.engine, NET_F, 'generator')
        self.buffer = self.engine.buffer                # cgan.py:65
        self.outdimsize = self.engine.outdimsize        # cgan.py:66
        self.is3d = is3d
        self.exp_name = exp_name
        self.meanstd_x, self.meanstd_y = meanstd_x, meanstd_y
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("distributed=True needs torch.distributed.init_process_group first")
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # checkpoints (cgan.py:84-103): own flat format, same directory convention, max_to_keep=50
        self.checkpoint_path = checkpoint_dir or f"./checkpoints/train_{exp_name}"
        self.max_to_keep = 50
        # Restore FIRST, then replicate: rank 0 alone resolves and reads the checkpoint (the directory need not be shared
        # and must not be listed at different moments by different ranks); tem_comm_sync_params then ships rank 0's
        # parameters and Adam moments to every replica and the step counter travels beside them.
        if self.rank == 0:
            if ckpt_restore is not None:
                self.restore(ckpt_restore)
                print(f"checkpoint {ckpt_restore} restored")
            else:
                latest = self.latest_checkpoint()
                if latest:
                    self.restore(latest)
                    print('Latest checkpoint restored!!')
        if distributed:
            self._init_distributed()

    # ---- data parallel (README.md:93-94, cgan.py:8-11) ----------------------------------------
    def _init_distributed(self):
        import torch.distributed as dist

        def bcast(raw):
            obj = [raw]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        self.engine.init_comm(self.rank, self.world, bcast)       # NCCL communicator + broadcast of params / m / v from rank 0
        self.engine.step = int(bcast(self.engine.step if self.rank == 0 else None))   # Adam bias correction uses the step

    # ---- checkpoints ----------------------------------------------------------------------------
    def _ckpts(self):
        fs = glob.glob(os.path.join(self.checkpoint_path, "ckpt-*.npz"))
        return sorted(fs, key=lambda f: int(os.path.basename(f)[5:-4]))

    def latest_checkpoint(self):
        fs = self._ckpts()
        return fs[-1] if fs else None

    def make_checkpoint(self, epoch_num):
        """cgan.py:105-107: saves the four networks and the four optimizers' state."""
        path = None
        if self.rank == 0:
            os.makedirs(self.checkpoint_path, exist_ok=True)
            fs = self._ckpts()
            nxt = (int(os.path.basename(fs[-1])[5:-4]) + 1) if fs else 1
            path = os.path.join(self.checkpoint_path, f"ckpt-{nxt}.npz")
            data = {"step": np.int64(self.engine.step), "wf": np.int64(self.engine.wf), "is3d": np.int64(self.is3d),
                    "dimsize": np.int64(self.engine.dimsize)}
            for name, net in (("generator_g", NET_G), ("generator_f", NET_F), ("discriminator_x", NET_DX), ("discriminator_y", NET_DY)):
                data[name] = self.engine.get_vector(net, 0)
                data[name + "_optimizer_m"] = self.engine.get_vector(net, 2)
                data[name + "_optimizer_v"] = self.engine.get_vector(net, 3)
            np.savez(path, **data)
            for old in self._ckpts()[:-self.max_to_keep]:
                os.remove(old)
        if self.world > 1:       # every rank returns the path rank 0 wrote, and only after it exists
            import torch.distributed as dist
            obj = [path]
            dist.broadcast_object_list(obj, src=0)
            path = obj[0]
            dist.barrier()
        print(f"Saving checkpoint for epoch {epoch_num} at {path}")
        return path

    def restore(self, path):
        z = np.load(path)
        if int(z["wf"]) != self.engine.wf or bool(z["is3d"]) != bool(self.is3d):
            raise RuntimeError("checkpoint does not match this model (wf / is3d)")
        for name, net in (("generator_g", NET_G), ("generator_f", NET_F), ("discriminator_x", NET_DX), ("discriminator_y", NET_DY)):
            self.engine.set_vector(net, z[name], 0)
            self.engine.set_vector(net, z[name + "_optimizer_m"], 2)
            self.engine.set_vector(net, z[name + "_optimizer_v"], 3)
        self.engine.step = int(z["step"])

    # ---- training -------------------------------------------------------------------------------
    def train_step(self, real_x, real_y):
        """cgan.py:144-230.  real_x / real_y: [B,n,n,n,1] float32 (standardised) or uint8 together with
        self.meanstd_x / self.meanstd_y (standardisation fused into the first-layer kernels).
        Returns (total_gen_g, total_gen_f, disc_y, disc_x, gen_g, gen_f, total_cycle)."""
        return self.engine.train_step(real_x, real_y, self.meanstd_x, self.meanstd_y)

    def train(self, train_input, train_target, epochs=3000, start=0, debug=False, sample=None, sample_gt=None,
              enable_eager=False, num_samples=4096, check_freq=1):
        """cgan.py:242-287.  train_input / train_target: iterables of batches (re-iterable per epoch)."""
        for epoch in range(start, start + epochs):
            t0 = time.time()
            loss = np.zeros((7), dtype=np.float32)
            count = 0
            for data_f, data_g in zip(train_input, train_target):
                loss += np.asarray(self.train_step(data_f, data_g), np.float32)
                count += 1
            loss = loss / max(count, 1)
            print(f"Epoch {epoch+1} loss [g_gen_total, f_gen_total, disc_y, disc_x, g_gen_only, f_gen_only, cycle]: {loss}")
            if (epoch + 1) % check_freq == 0:
                self.make_checkpoint(epoch + 1)
                if debug and sample is not None:
                    sample_pred = self.predict(sample)
                    if sample_gt is not None:
                        b = self.buffer
                        gt = np.asarray(sample_gt)
                        gt = gt[:, b:-b, b:-b, b:-b, :] if self.is3d else gt[:, b:-b, b:-b, :]
                        pred = sample_pred.cpu().numpy() if isinstance(sample_pred, torch.Tensor) else sample_pred
                        print(f"Accuracy on sample: {float(np.sqrt(np.mean((gt[0] - pred[0]) ** 2)))}")   # debug.py:65-71
            print(f"Time taken for epoch {epoch+1} is {time.time()-t0}")

    def predict(self, data):
        """cgan.py:289-293: generator_g in inference mode (no dropout)."""
        return self.engine.gen_forward(NET_G, data, meanstd=self.meanstd_x if _is_u8(data) else None)


def _is_u8(x):
    return getattr(x, "dtype", None) in (np.uint8, torch.uint8)
