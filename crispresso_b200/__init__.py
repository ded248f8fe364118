"""crispresso_b200 -- B200-native (sm_100a CUDA behind a C ABI) implementation of CRISPResso's
read->amplicon alignment + indel quantification hot path (CRISPRessoCORE.py:1547-1583,
1791-2072, 2773-2864).  See DESIGN.md and INTEGRATION.md."""
from ._lib import Context, CrgpuError, load  # noqa: F401
