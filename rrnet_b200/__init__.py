"""rrnet_b200 -- B200-native (sm_100a) implementation of RRNet's post-backbone detection hot path.

    rrnet_b200.ops      torch-tensor front end of the C ABI (include/rrnet_b200.h)
    rrnet_b200.host     host-side mirror of the reference's Python interface for this path
    rrnet_b200.synth    seeded synthetic inputs
    rrnet_b200.build    nvcc build of librrnet_b200.so

Importing the package does not load the CUDA library; the first op call does, and raises if the
library is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"
