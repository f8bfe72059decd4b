"""Import helper: the package directory is named `miden-vm_b200` (not a valid Python identifier),
so it is loaded under the module name `miden_vm_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def load_pkg():
    if "miden_vm_b200" in sys.modules:
        return sys.modules["miden_vm_b200"]
    d = os.path.join(ROOT, "miden-vm_b200")
    spec = importlib.util.spec_from_file_location(
        "miden_vm_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["miden_vm_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
