"""GeGLU / SwiGLU parameter holders (reference: rails/similarities/layers.py:19-74).

Same parameter names (`_w` (in, 2*out), `_b` (1, 2*out)) and the same initialisers, so reference
checkpoints load unchanged.  The arithmetic `act(xW_a+b_a) * (xW_b+b_b)` runs inside the CUDA query
prologue (rails_b200/csrc/mol_prologue.cu: glu_kernel); these modules hold no torch math.
"""
import torch


class _GLU(torch.nn.Module):
    kind = -1

    def __init__(self, in_features: int, out_features: int) -> None:
        super().__init__()
        self._in_features = in_features
        self._out_features = out_features
        self._w = torch.nn.Parameter(torch.empty((in_features, out_features * 2)).normal_(mean=0, std=0.02))
        self._b = torch.nn.Parameter(torch.zeros((1, out_features * 2)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # pragma: no cover
        raise RuntimeError(
            "rails_b200 layers are parameter holders; the GLU runs inside the CUDA query prologue "
            "(use MoLSimilarity / MoLBruteForceTopK)"
        )


class GeGLU(_GLU):
    kind = 0


class SwiGLU(_GLU):
    kind = 1
