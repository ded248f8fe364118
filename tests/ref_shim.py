"""Import the UNMODIFIED reference (CRISPResso/CRISPRessoCORE.py) in the build container.

The reference checks for `java`, `flash`, `needle` on PATH and imports matplotlib / seaborn /
Biopython at import time (CORE:31-36, 349-368); none of them is needed by process_df_chunk.
This shim puts three dummy executables on PATH and stub modules in sys.modules, then imports the
reference from /root/reference.  Used only by tests/golden/make_golden.py and
by tests that are skipped when /root/reference is absent (the GPU box).
"""
import os
import stat
import sys
import tempfile
import types
from unittest import mock

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CRISPResso"))


_core = None


def load_core():
    global _core
    if _core is not None:
        return _core
    if not available():
        raise RuntimeError("reference tree not present")
    d = tempfile.mkdtemp(prefix="crref_bin_")
    for exe in ("java", "flash", "needle"):
        p = os.path.join(d, exe)
        with open(p, "w") as f:
            f.write("#!/bin/sh\nexit 0\n")
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    os.environ["PATH"] = d + os.pathsep + os.environ.get("PATH", "")
    for name in ("matplotlib", "matplotlib.backends", "matplotlib.backends.backend_pdf", "matplotlib.font_manager",
                 "matplotlib.colors", "matplotlib.gridspec", "matplotlib.lines", "matplotlib.pyplot", "matplotlib.cm",
                 "matplotlib.patches", "pylab", "seaborn", "seaborn.matrix", "Bio", "Bio.SeqIO", "Bio.pairwise2",
                 "Bio.Seq", "Bio.SeqRecord"):
        if name not in sys.modules:
            m = mock.MagicMock(name=name)
            m.__path__ = []
            sys.modules[name] = m
    hm = types.ModuleType("seaborn.matrix")

    class _HeatMapper(object):
        pass

    hm._HeatMapper = _HeatMapper
    sys.modules["seaborn.matrix"] = hm
    sys.modules["seaborn"].matrix = hm
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import CRISPResso.CRISPRessoCORE as core
    _core = core
    return core
