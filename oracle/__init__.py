"""oracle/ -- TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's cpu_baseline).

Nothing under pwc_net_pytorch_b200/ imports this package; the product path has no CPU fallback.

  pwc_oracle.c     plain-C restatement of the reference's warp + correlation path
  c_oracle.py      ctypes/numpy wrapper around it (build: oracle/build_oracle.sh)
  torch_ref.py     closed-form PyTorch restatement (fp32/fp64, autograd) + the port of the
                   reference's PyTorch-level path (WarpingLayer + CostVolumeLayer) used as the
                   CPU timing baseline
  build_ref.sh     compiles the reference's own correlation_cuda_kernel.cu, unchanged, for
                   sm_100a into oracle/_ref/libref_corr.so (GPU-side parity pin + GPU reference bar)
  ref_cuda.py      ctypes driver for oracle/_ref/libref_corr.so (restates correlation_cuda.c)
"""
