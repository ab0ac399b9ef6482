"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

The reference (gseg-ethz/fusion4landslide, mounted read-only at /root/reference in
the authoring container) imports Open3D, hnswlib, faiss, easydict, SWIG builds ...
none of which exist in this image.  Its pure torch/numpy functions on the hot path
(src/functions.py, scripts/weighted_svd.py, src/models/outlier_classifier.py) do
not need them, so we register empty stand-in modules for the missing imports and
import the reference files from where they lie.  Nothing is copied.

Used only by oracle/make_golden.py (to pin the oracle and to write tests/golden/*)
and by tests that are skipped when /root/reference is absent (the GPU box).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("F4L_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "open3d", "hnswlib", "faiss", "easydict", "laspy", "coloredlogs", "colorhash",
    "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors", "hydra",
    "cpp_core", "cpp_core.pcd_tiling", "cpp_core.pcd_tiling.build",
    "cpp_core.pcd_tiling.build.pcd_tiling",
    "cpp_core.supervoxel_segmentation", "cpp_core.supervoxel_segmentation.build",
    "cpp_core.supervoxel_segmentation.build.supervoxel",
    "superpoint_transformer", "superpoint_transformer.src",
]


class _Anything(types.ModuleType):
    """Module whose every attribute is another permissive stub (import-time only)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")


class _StubFinder:
    """Last-resort import hook: any submodule of a package that is absent from this image (Open3D, the
    superpoint_transformer submodule, image matchers ...) resolves to a permissive stand-in, so that
    src/coarse_to_fine_matching_base.py can be imported for its module-level functions."""
    roots = ("superpoint_transformer", "open3d", "cpp_core", "matplotlib", "romatch", "hydra", "omegaconf", "kornia",
             "cv2", "pycolmap", "pyproj", "laspy", "plyfile", "hnswlib", "faiss", "easydict", "coloredlogs", "colorhash")

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        if name.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Anything(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def load():
    """Make `import src.functions`, `import scripts.weighted_svd` ... resolve to the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())          # before the loop: stubbed packages must be PACKAGES
    for name in _STUBS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if "easydict" in sys.modules and isinstance(sys.modules["easydict"], _Anything):
        class EasyDict(dict):
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__
        sys.modules["easydict"].EasyDict = EasyDict
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def ref_base():
    """src/coarse_to_fine_matching_base.py (module-level functions: map_corr_2d_to_3d, ...)."""
    load()
    return importlib.import_module("src.coarse_to_fine_matching_base")


def ref_functions():
    load()
    return importlib.import_module("src.functions")


def ref_weighted_svd():
    load()
    return importlib.import_module("scripts.weighted_svd")


def ref_outlier_classifier():
    load()
    return importlib.import_module("src.models.outlier_classifier")
