"""Import alias: ``import b2048`` loads the package directory
``rl-2048-with-reinforce-and-actor-critic_b200/`` (whose name is not a valid Python identifier)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rl-2048-with-reinforce-and-actor-critic_b200")
_spec = importlib.util.spec_from_file_location("b2048", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b2048"] = _mod
_spec.loader.exec_module(_mod)
