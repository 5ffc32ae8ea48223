from torchflows_b200.bijections.finite.autoregressive.architectures import (NICE, RealNVP, InverseRealNVP, MAF, IAF,
                                                                           CouplingRQNSF, MaskedAutoregressiveRQNSF,
                                                                           InverseAutoregressiveRQNSF)
