"""Drop-in module name for the reference's cross_atten/pscan.py (imported at mamba.py:9)."""
from gfe_mamba_b200.pscan import PScan, npo2, pad_npo2, pscan  # noqa: F401
