"""Minimal stand-in for the `diffusers` symbols that /root/reference/hyvideo/vae imports.

TEST INFRASTRUCTURE ONLY. The real `diffusers` (pinned 0.31.0 in the reference's
requirements.txt:2) is not installed in this image and there is no network, so this shim
restates just enough of its published behaviour for the UNMODIFIED reference VAE files to
import and run on CPU.  It is used by oracle/make_golden.py (here, where /root/reference
exists) to produce the committed fixtures under tests/golden/.  Nothing in the product
package imports it.
"""
__version__ = "0.31.0-shim"
