"""distributions_b200 -- B200-native batched mixture scoring + sampling (one hot path of
forcedotcom/distributions), behind the C-ABI of include/dist_b200.h.

  distributions_b200.capi     ctypes binding of libdist_b200.so (raises if the library is missing)
  distributions_b200.mixture  host-side mirror of the reference's Shared / Group / Mixture interface
  distributions_b200.synth    seeded synthetic workloads (numpy)
  distributions_b200.build    in-tree nvcc build of the library
"""
__version__ = "0.1.0"
