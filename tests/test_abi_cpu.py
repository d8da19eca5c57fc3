"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/omnisafe_b200.h declares (no compute calls: there is no GPU here and no CPU fallback)."""
import ctypes
import os

from omnisafe_b200 import _lib


def test_header_symbols_are_exported():
    protos = _lib.parse_header()
    assert len(protos) >= 10
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(dll, name), f'{name} declared in include/omnisafe_b200.h but not exported'


def test_library_loads_and_reports_version():
    L = _lib.lib()
    assert L.osb_abi_version() == 1
    assert os.path.basename(_lib.LIB_PATH) == 'libomnisafe_b200.so'


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.abspath(_lib.__file__))
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f'{f} imports the oracle'
