"""torch.hub entry points. The reference's hubconf.py:1-5 is a DeiT leftover (``from models import *`` fails: there is
no ``models`` package); this one exports the PVLT entry points it was meant to expose."""
dependencies = ["torch"]

from mvlt_b200.libs.pvlt import pvlt_large, pvlt_medium, pvlt_small, pvlt_tiny  # noqa: E402,F401
