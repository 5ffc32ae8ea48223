"""torchflows_b200 -- B200-native (sm_100a) implementation of the coupling / masked-autoregressive hot path of
torchflows: ``Flow.log_prob / sample / fit``, the presets RealNVP, NICE, MAF, IAF, CouplingRQNSF,
MaskedAutoregressiveRQNSF and the ``Bijection.forward / inverse -> (z, log_det)`` contract, with the
reference's module paths and ``state_dict`` layout.  Compute goes through hand-written CUDA kernels behind a
C ABI (include/b2f.h, torchflows_b200/lib/libb2f.so); there is no CPU fallback."""
from torchflows_b200._version import __version__
from torchflows_b200.flows import Flow
from torchflows_b200._program import set_math_mode

__all__ = ['Flow', 'set_math_mode', '__version__']
