"""miden-vm_b200: Blackwell-native STARK proving backend for Miden VM.

The product is the CUDA shared library `csrc/libmiden_b200.so` (C ABI in include/miden_b200.h),
which contains the sm_100a kernels *and* the C++ host orchestration that mirrors
`ProverInstance::prove` (reference crates/lifted-stark/src/prover/mod.rs:230-578).  The Python
modules here are thin host-side helpers used by tests and bench.py:

  binding      ctypes view of the C ABI (fails loudly if the library is missing)
  air_program  constructor for the AIR constraint op-list (`mdn_air.program`)
  workload     Miden production configuration + synthetic trace generator
"""
from . import air_program, binding, parallel, workload  # noqa: F401
