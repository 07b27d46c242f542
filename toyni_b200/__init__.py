"""toyni_b200 — B200-native (sm_100a) BabyBear prover hot path behind toyni's C ABI.

Host-side mirror of the reference's interfaces for this path:
  ntt.py     <- src/ntt.rs        (ntt_cuda, intt_cuda, cuda_available, CudaBuffer)
  domain.py  <- src/math/domain.rs (BabyBearDomain.fft / ifft / fft_ext / ifft_ext)
  fri.py     <- src/math/fri.rs    (fri_fold, fri_fold_ext) and the prover's FRI commit loop
  merkle.py  <- src/merkle.rs + src/fibonacci.rs:340-374 (salted / unsalted trees, openings)
  device.py  — device-resident (torch tensor) forms used by the bench and the multi-GPU layer
  multigpu.py — sharded NTT / fold chain over torch.distributed (NCCL)
All compute goes through libntt_cuda.so (toyni_b200/csrc); there is no CPU fallback.
"""
from .lib import P, library_path  # noqa: F401
