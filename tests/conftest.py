import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests', 'hostemu')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    # compiled specialised kernels: keep the tests' cubins out of ~/.cache
    if 'AHK_JIT_CACHE' not in os.environ:
        import tempfile
        os.environ['AHK_JIT_CACHE'] = tempfile.mkdtemp(prefix='ahk_jit_cache_')
    # fresh checkout: build the CUDA library (nvcc cross-compiles without a GPU), the
    # oracle and the host-emulation library once, like __graft_entry__.build() does
    import subprocess
    for lib, script in ((os.path.join(ROOT, 'ahkab_b200', 'lib', 'libahkab_b200.so'),
                         os.path.join(ROOT, 'ahkab_b200', 'csrc', 'build.sh')),
                        (os.path.join(ROOT, 'oracle', '_build', 'libahk_oracle.so'),
                         os.path.join(ROOT, 'oracle', 'build.sh')),
                        (os.path.join(ROOT, 'tests', 'hostemu', '_build', 'libahk_hostemu.so'),
                         os.path.join(ROOT, 'tests', 'hostemu', 'build.sh'))):
        if not os.path.exists(lib):
            subprocess.check_call(['sh', script], stdout=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    """A machine without a CUDA device skips the GPU tests instead of failing them (the engine
    itself still fails loudly there: tests/test_abi.py::test_no_cpu_fallback)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200)")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


GOLD = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def gold():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLD, name + '.npz'))
    return load
