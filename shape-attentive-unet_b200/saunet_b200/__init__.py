"""saunet_b200 -- B200-native execution engine behind the reference's nn.Module surface.

  _C        ctypes binding of libsaunet_b200.so (include/saunet_b200.h); raises if the library is missing
  engine    NHWC buffers, the backward tape, thin op wrappers, the autograd boundary
  blocks    fwd + hand-written bwd of every SAUNet block
  parallel  one-process-per-GPU data parallelism: flat gradient arena + NCCL all-reduce
  synth     synthetic ACDC-shaped inputs and deterministic weights (host side)
"""
