"""occ_b200 — B200-native (sm_100a) DFT exchange-correlation build behind OCC's occ::dft::DFT interface.

Only the hot path (DFT::compute_K_dft, /root/reference include/occ/dft/dft.h:190-388) lives here:
  csrc/      hand-written CUDA kernels + the C-ABI shared library (include/occ_xc_b200.h)
  aobasis.py host-side mirror of occ::gto::AOBasis (basis JSON loading, shell normalisation)
  molecules.py  geometries / synthetic density matrices of the BASELINE configs
  dft.py     host-side mirror of occ::dft::DFT over the C ABI (ctypes; PyTorch only for device buffers)
The product path fails loudly if the CUDA library is missing; there is no CPU fallback.
"""
__version__ = "0.1.0"
