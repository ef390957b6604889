"""Import shim: `import vaegan_b200` loads the package that lives in the (hyphenated, hence not directly importable)
directory `vae-gan-based-model-for-image-generation-and-denoising_b200/`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "vae-gan-based-model-for-image-generation-and-denoising_b200")
_spec = importlib.util.spec_from_file_location("vaegan_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["vaegan_b200"] = _mod
_spec.loader.exec_module(_mod)
