"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/lcr_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def library():
    import lcrnet_b200.build as b
    return ctypes.CDLL(b.build_library())


def _declared():
    text = open(os.path.join(REPO, 'include', 'lcr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lcr_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(library):
    names = _declared()
    assert len(names) >= 6
    for n in names:
        assert hasattr(library, n), 'missing symbol %s' % n


def test_binding_table_matches_header(library):
    from lcrnet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    assert _lib.lib().lcr_abi_version() >= 1


def test_no_cpu_fallback():
    import torch
    from lcrnet_b200 import ext
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        ext.grid_subsampling(torch.zeros(4, 3), torch.tensor([4]), 0.5)
    with pytest.raises(RuntimeError):
        ext.radius_neighbors(torch.zeros(4, 3), torch.zeros(4, 3), torch.tensor([4]), torch.tensor([4]), 1.0)


def test_argument_checks_like_reference():
    import torch
    from lcrnet_b200 import ext
    with pytest.raises(RuntimeError, match='float'):
        ext.grid_subsampling(torch.zeros(4, 3, dtype=torch.float64), torch.tensor([4]), 0.5)
    with pytest.raises(RuntimeError, match='long'):
        ext.grid_subsampling(torch.zeros(4, 3), torch.tensor([4], dtype=torch.int32), 0.5)
    with pytest.raises(RuntimeError, match='contiguous'):
        ext.radius_neighbors(torch.zeros(3, 4).t(), torch.zeros(4, 3), torch.tensor([4]), torch.tensor([4]), 1.0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, 'lcr-net_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'lcr_oracle' not in text, f
