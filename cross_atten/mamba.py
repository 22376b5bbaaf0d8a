"""Drop-in module name: ``from cross_atten.mamba import Mamba, MambaConfig, MambaBlock, RMSNorm`` (the imports of
the reference's mamba_transformer.py:8 and jamba.py:9) resolve here when this repository precedes the reference
on sys.path.  ``cross_atten`` is a namespace package in both trees, so the reference's other modules
(mamba_transformer, jamba, ...) keep resolving to the reference."""
from gfe_mamba_b200.mamba import Mamba, MambaBlock, MambaConfig, ResidualBlock, RMSNorm  # noqa: F401
from gfe_mamba_b200.pscan import pscan  # noqa: F401
