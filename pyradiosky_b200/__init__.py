"""pyradiosky_b200 -- B200-native (sm_100a) implementation of pyradiosky's per-component hot path.

Public surface (mirrors the reference, src/pyradiosky/__init__.py):
    SkyModel                 update_positions, coherency_calc, at_frequencies, jansky_to_kelvin, ...
    utils.stokes_to_coherency, utils.jy_to_ksr
    astrometry.Time, astrometry.EarthLocation   (stand-ins for the astropy types the reference takes)
    astrometry.IERSTable                        (UT1-UTC and polar motion from an IERS finals2000A file)

Importing the package does not need a GPU; constructing a SkyModel or calling a kernel does, and the
built ``csrc/librsk.so`` must exist (``__graft_entry__.build()``).  There is no CPU fallback.
"""

from . import astrometry, utils  # noqa: F401
from .astrometry import EarthLocation, IERSTable, Time  # noqa: F401
from .skymodel import SkyModel, StepGraph  # noqa: F401

__version__ = "0.1.0"
