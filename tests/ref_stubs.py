"""Import stubs for packages the reference imports for code that is NOT on the hot path (trimesh, matplotlib, pylab:
datasets/data_utils.py:4,15, misc/visualize/vis_utils.py:1-10) and that are not installed in this image."""
import importlib.abc
import importlib.machinery
import sys
import types
from unittest import mock


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return mock.MagicMock()


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    P = ("trimesh", "matplotlib", "mpl_toolkits", "pylab")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.P:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
