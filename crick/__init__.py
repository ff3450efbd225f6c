"""Drop-in shim: ``import crick`` resolves to the B200-native build (crick_b200), with the
reference's export list (crick/__init__.py:1-16) and its submodule paths -- pickles made by
stock crick name ``crick.tdigest.TDigest`` etc.."""
from crick_b200 import (SpaceSaving, SummaryStats, TDigest, TDigestGroup, __git_revision__,
                        __numpy_version__, __version__, grouped_update)

from . import space_saving, stats, tdigest  # noqa: F401
