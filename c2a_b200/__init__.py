"""c2a_b200 -- B200-native (sm_100a CUDA, FP64) implementation of the controlled
conservative-advancement CCD hot path of EwhaGlab/C2A behind the reference's own API.

The compute lives in ``c2a_b200/csrc`` (CUDA kernels + the C ABI declared in ``include/c2a_b200.h``);
this Python package is plumbing for tests and the benchmark: ctypes bindings (``api``), mesh
inputs (``meshes``) and synthetic pose batches (``workloads``)."""
