"""Deterministic synthetic weights / inputs: re-exported from the product package's
hunyuanvideo_efficiency_b200/synthetic.py (a pure parameter/input generator with no VAE arithmetic), so that
the reference (through load_state_dict), the oracle and the CUDA path are given bit-identical parameters."""
from hunyuanvideo_efficiency_b200.synthetic import (HY_VAE_CONFIG, SMALL_CONFIG, make_latent, make_state_dict,  # noqa: F401
                                                    make_video, state_dict_spec)
