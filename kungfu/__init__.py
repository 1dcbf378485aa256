"""`kungfu` — import-name compatibility with the reference: `import kungfu.torch as kf`, `kungfu.python`, `kungfu.cmd`, and the
`kungfu.tensorflow.*` module paths, which resolve to the implementations of the same names that work on torch tensors
(`kungfu.tensorflow.optimizers.SynchronousSGDOptimizer` is the torch optimizer wrapper, `kungfu.tensorflow.ops.all_reduce` the
collective on a torch tensor, …). Everything lives in `kungfu_b200`; a meta-path finder maps `kungfu.X[.Y]` onto it on demand, so a
module that exists there is importable under the reference's path without being listed here.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys

import kungfu_b200 as _impl

__version__ = _impl.__version__

# reference path (below `kungfu.`) -> implementation path (below `kungfu_b200.`); longest prefix wins, the rest of the dotted name is
# appended when it exists there and dropped otherwise (`kungfu.tensorflow.optimizers.keras` -> the optimizers package)
_PREFIXES = {
    "_utils": "utils",
    "tensorflow": "torch",
    "tensorflow.ops": "ops",
    "tensorflow.ops.local": "ops",
    "tensorflow.ops._tf_oplib": "ops",
    "tensorflow.optimizers": "torch.optimizers",
    "tensorflow.initializer": "initializer",
    "tensorflow.policy": "elastic.policy",
    "tensorflow.hooks": "elastic.hooks",
    "tensorflow.experimental.hook": "elastic.hooks",
    "tensorflow.experimental": "elastic",
    "tensorflow.variables": "variables",
    "tensorflow.compat": "torch",
    "tensorflow.v1": "torch",
    "tensorflow.v2": "torch",
    "tensorflow.v1.helpers": "helpers",
    "tensorflow.v1.datasets": "elastic.datasets",
    "tensorflow.v1.benchmarks": "benchmarks",
    "tensorflow.v1.examples": "examples",
    "tensorflow.v2.examples": "examples",
}


def _resolve(fullname: str):
    rel = fullname[len("kungfu."):]
    parts = rel.split(".")
    for k in range(len(parts), 0, -1):
        head = ".".join(parts[:k])
        if head in _PREFIXES:
            base, rest = "kungfu_b200." + _PREFIXES[head], parts[k:]
            break
    else:
        base, rest = "kungfu_b200", parts
    while True:
        name = ".".join([base] + rest)
        try:
            if importlib.util.find_spec(name) is not None:
                return name
        except (ImportError, AttributeError, ValueError):
            pass
        if not rest or base == "kungfu_b200":
            return None
        rest = rest[:-1]  # a leaf the implementation keeps in its parent module


class _MainSpec(importlib.machinery.ModuleSpec):
    real_parent = ""

    @property
    def parent(self):
        return self.real_parent


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.InspectLoader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("kungfu."):
            return None
        real = _resolve(fullname)
        if real is None:
            return None
        spec = importlib.util.find_spec(real)
        if fullname.endswith(".__main__"):
            # `python -m kungfu.x`: the implementation's __main__ runs inside ITS package, so that its relative imports resolve there
            main = _MainSpec(fullname, self, origin=spec.origin)
            main.real_parent = real.rpartition(".")[0]
            return main
        # every alias is importable as a parent (`kungfu.tensorflow.initializer.keras` hangs below what is a plain module here): the
        # import system gives a plain-module target an empty __path__, under which only this finder resolves names
        return importlib.util.spec_from_loader(fullname, self, is_package=True)

    def create_module(self, spec):
        module = importlib.import_module(_resolve(spec.name))  # the implementation module itself, under a second name
        self._real_spec[spec.name] = module.__spec__
        return module

    def exec_module(self, module):
        # the import system has just stamped the alias spec on the module; the module keeps its own identity (`__spec__.parent` is what
        # its relative imports resolve against)
        for alias, real in list(self._real_spec.items()):
            if real is not None and real.name == module.__name__:
                module.__spec__ = real
                del self._real_spec[alias]

    _real_spec = {}

    # `python -m kungfu.info`, `python -m kungfu.tensorflow.v1.examples`: runpy asks the loader for the code of `<package>.__main__`
    def get_code(self, fullname):
        real = _resolve(fullname)
        return importlib.util.find_spec(real).loader.get_code(real)

    def get_source(self, fullname):
        real = _resolve(fullname)
        return importlib.util.find_spec(real).loader.get_source(real)

    def is_package(self, fullname):
        return importlib.util.find_spec(_resolve(fullname)).submodule_search_locations is not None


if not any(isinstance(f, _Alias) for f in sys.meta_path):
    sys.meta_path.insert(0, _Alias())  # ahead of the path finder: a parent's __path__ must never load a SECOND copy of a submodule

for _name in ("python", "torch", "cmd", "ops", "elastic", "initializer", "checkpoint", "utils", "models", "parallel", "helpers", "platforms",
              "benchmarks", "variables", "examples", "info", "loader"):
    globals()[_name] = importlib.import_module(f"kungfu.{_name}")
_utils = importlib.import_module("kungfu._utils")  # reference: kungfu/_utils.py (measure, show_duration, _log_event)
tensorflow = importlib.import_module("kungfu.tensorflow")
