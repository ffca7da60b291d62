"""mira_b200 — B200-native commitment (MSM) engine behind Mira's CommitmentKey::commit.

Only what the hot path needs: csrc/ (CUDA kernels + C ABI -> libmira_b200.so) and the host-side
mirror of the reference interface (commitment.CommitmentKey)."""
from .commitment import (BN254_G1, GRUMPKIN_G1, IDENTITY, CommitmentKey, CudaError, NotOnCurve, TooLongInput,
                         combine_partials, combine_partials_device)

__all__ = ["BN254_G1", "GRUMPKIN_G1", "IDENTITY", "CommitmentKey", "CudaError", "NotOnCurve", "TooLongInput",
           "combine_partials", "combine_partials_device"]
