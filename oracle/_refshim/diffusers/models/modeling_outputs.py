from dataclasses import dataclass
from typing import Any, Optional


@dataclass
class AutoencoderKLOutput:
    # the reference passes tiles_ci=None at autoencoder_kl_causal_3d.py:296
    latent_dist: Any = None
    tiles_ci: Optional[Any] = None
