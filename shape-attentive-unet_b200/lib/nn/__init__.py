"""Mirror of the one name the hot path imports from the reference's lib/nn:
``SynchronizedBatchNorm2d`` (lib/nn/modules/batchnorm.py:30-61).

The reference's thread-based DataParallel / sync-BN rendezvous machinery is NOT
rebuilt (out of scope; replaced by one process per GPU
+ NCCL gradient all-reduce, see saunet_b200/parallel.py).  Outside DataParallel
the reference layer IS ``F.batch_norm`` with momentum 0.001 plus three unused
buffers, which is what this class reproduces.
"""
import torch
import torch.nn as nn


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.001, affine=True):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        # lib/nn/modules/batchnorm.py:50-54: present in the state_dict, untouched outside DataParallel
        self.register_buffer("_tmp_running_mean", torch.zeros(num_features))
        self.register_buffer("_tmp_running_var", torch.ones(num_features))
        self.register_buffer("_running_iter", torch.ones(1))
