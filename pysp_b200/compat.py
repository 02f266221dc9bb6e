"""Optional alias: make this package importable under the reference's package name `pySP`.

pySP is meant to be vendored as a sub-module and uses absolute imports such as
`from pySP.colorize.transform import cam_to_lin_srgb` (README.md:32; base_types/image_base.py:7-10).  After

    import pysp_b200.compat
    pysp_b200.compat.install_as_pySP()

every `pySP` / `pySP.<sub.module>` import resolves to `pysp_b200` / `pysp_b200.<sub.module>` (the SAME module objects, not
copies), for the modules the B200 path mirrors: image, const, normalization, raw_hdr, raw_correction, raw_bad_pixel_corr,
base_types.image_base, colorize.transform, colorize.rgb_space, debayer (debayer_ahd / debayer_eag), wb_cct.cam_wb,
wb_cct.helpers_cam_mat, dng_warp_corr.*.  It is opt-in because a process can only have one `pySP`: it refuses to shadow a
real pySP that is already imported.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

ALIAS = "pySP"
TARGET = __name__.rsplit(".", 1)[0]          # "pysp_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)      # the module object itself: `pySP.x is pysp_b200.x`

    def exec_module(self, module):
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != ALIAS and not fullname.startswith(ALIAS + "."):
            return None
        real = TARGET + fullname[len(ALIAS):]
        try:
            mod = importlib.import_module(real)
        except ImportError:
            return None
        spec = importlib.machinery.ModuleSpec(fullname, _AliasLoader(real), is_package=hasattr(mod, "__path__"))
        return spec


_finder = None


def install_as_pySP():
    """Route `import pySP[...]` to this package.  Raises if a different `pySP` is already imported."""
    global _finder
    existing = sys.modules.get(ALIAS)
    if existing is not None and existing is not sys.modules.get(TARGET):
        raise ImportError("a different `pySP` package is already imported (%r); pysp_b200 will not shadow it" % (
            getattr(existing, "__file__", existing),))
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)


def uninstall():
    """Remove the alias and every `pySP*` entry it created from sys.modules."""
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in [n for n in sys.modules if n == ALIAS or n.startswith(ALIAS + ".")]:
        mod = sys.modules[name]
        if getattr(mod, "__name__", "").startswith(TARGET):
            del sys.modules[name]
