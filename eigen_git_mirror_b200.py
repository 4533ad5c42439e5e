"""Import alias: the product package lives in the directory ``eigen-git-mirror_b200/`` (the name the project
layout prescribes), which is not a valid Python identifier.  ``import eigen_git_mirror_b200`` loads that directory
as a regular package under this name, so ``from eigen_git_mirror_b200 import solvers`` works as usual."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eigen-git-mirror_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
