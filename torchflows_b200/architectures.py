"""``from torchflows_b200.architectures import RealNVP`` -- the import path the reference's README documents
(README.md:13-14); the presets live in bijections/finite/autoregressive/architectures.py."""
from torchflows_b200.bijections.finite.autoregressive.architectures import *  # noqa: F401,F403
from torchflows_b200.bijections.finite.autoregressive.architectures import (  # noqa: F401
    AutoregressiveArchitecture, NICE, RealNVP, InverseRealNVP, MAF, IAF, CouplingRQNSF, MaskedAutoregressiveRQNSF,
    InverseAutoregressiveRQNSF, CouplingLRS, MaskedAutoregressiveLRS, InverseAutoregressiveLRS)
