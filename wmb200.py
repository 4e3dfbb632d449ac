"""Import alias: the package directory name (mandated by the repo layout) contains hyphens,
so `import wmb200` loads it from its directory under a valid module name."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "audio-watermarking-deep-learning-watermarks-for-authenticating-speech_b200")
_spec = importlib.util.spec_from_file_location("wmb200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["wmb200"] = _mod
_spec.loader.exec_module(_mod)
