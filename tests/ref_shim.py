"""Import the UNMODIFIED reference (CRISPResso/CRISPRessoCORE.py) in the build container.

The reference checks for `java`, `flash`, `needle` on PATH and imports matplotlib / seaborn /
Biopython at import time (CORE:31-36, 349-368); none of them is needed by process_df_chunk.
This shim puts three dummy executables on PATH and stub modules in sys.modules, then imports the
reference from /root/reference -- or from its installed copy baseline/_ref -- (SURVEY.md App. E).  Used by
tests/golden/make_golden.py, by tests that are skipped when neither exists, and by the CPU baseline of bench.py
(oracle/ref_quantify.py).
"""
import os
import stat
import sys
import tempfile
import types
from unittest import mock

# the reference tree of the build container, else the copy `pip install --target baseline/_ref /root/reference` left in the
# repo (git-ignored; it travels to the GPU box, where bench.py's CPU baseline times the reference's own process_df_chunk)
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "CRISPResso")):
    REFERENCE_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


class PlotStub(object):
    """Stands in for pylab / matplotlib objects so that run_crispresso (CORE:1216) runs to its return statement
    without matplotlib: every attribute is a PlotStub, every call returns one, except the few calls whose return
    value the reference computes with: pylab.histogram IS numpy.histogram (CORE:2357-2365), pie /
    stem unpack into three values (CORE:2177, 3248), tick and limit getters return numbers (CORE:888, 1122, 2391,
    3456)."""

    def __init__(self, name=""):
        object.__setattr__(self, "_name", name)

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        if name == "histogram":
            import numpy
            return numpy.histogram
        return PlotStub(name)

    def __setattr__(self, name, value):
        pass

    def __call__(self, *args, **kwargs):
        n = self._name
        if n in ("pie", "stem"):
            return PlotStub(), [], []
        if n in ("get_view_interval", "get_ylim", "get_xlim"):
            return (0.0, 1.0)
        if n in ("get_yticks", "get_xticks"):
            return [0.0, 1.0]
        if n == "subplots":
            return PlotStub(), PlotStub()
        return PlotStub()

    def __getitem__(self, key):
        return PlotStub()

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CRISPResso"))


_core = None


def load_core(rich_plots=False):
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
    sys.modules["pylab"] = PlotStub("pylab")

    def globalxx(a, b):
        """Stand-in for Bio.pairwise2.align.globalxx in the reference's one-off HDR-amplicon sanity check
        (CORE:1367-1383, outside the hot path): an ungapped pairing, padded at the end."""
        n = max(len(a), len(b))
        return [(a.ljust(n, "-"), b.ljust(n, "-"), 0.0, 0, n)]

    sys.modules["Bio.pairwise2"].align.globalxx = globalxx
    sys.modules["Bio"].pairwise2 = sys.modules["Bio.pairwise2"]
    hm = types.ModuleType("seaborn.matrix")

    class _HeatMapper(object):
        """Base class of the reference's Custom_HeatMapper (CORE:840-935): keeps what its plot() reads."""

        def __init__(self, data, vmin, vmax, cmap, center, robust, annot, fmt, annot_kws, cbar, cbar_kws,
                     xticklabels=True, yticklabels=True, mask=None):
            import numpy
            self.data = self.plot_data = numpy.asarray(data)
            self.vmin, self.vmax, self.cmap, self.fmt = vmin, vmax, cmap, fmt
            self.annot, self.annot_data, self.annot_kws = False, numpy.zeros(0), annot_kws or {}
            self.xticks = self.yticks = self.xticklabels = self.yticklabels = []
            self.xlabel = self.ylabel = ""

    hm._HeatMapper = _HeatMapper
    sys.modules["seaborn.matrix"] = hm
    sys.modules["seaborn"].matrix = hm
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import CRISPResso.CRISPRessoCORE as core
    _core = core
    return core
