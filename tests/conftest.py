import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_addoption(parser):
    parser.addoption('--hostsim', action='store_true', default=False,
                     help='TEST INFRASTRUCTURE: run against the C-ABI library compiled for the HOST (tests/hostsim/: kernels on a '
                          'SIMT stand-in, no GPU) instead of sdf_b200/libsdfb200.so -- how tests/test_hostsim_library_cpu.py '
                          'runs a selection of the -m gpu tests where there is no GPU')


def pytest_configure(config):
    if config.getoption('--hostsim'):
        import util_hostsim
        from sdf_b200 import _abi
        _abi.LIB_PATH = util_hostsim.build_library()            # before the first _abi.lib() call of this process
        # run-time specialisation: the interpreter kernels unless a test asks (SDFB200_JIT=2); then the library's own
        # generated source is compiled for the host by a stand-in NVRTC (tests/hostsim/fake_nvrtc.cpp)
        os.environ.setdefault('SDFB200_JIT', '0')
        nvrtc, inc = util_hostsim.build_fake_nvrtc()
        os.environ['HOSTSIM_NVRTC'] = nvrtc
        os.environ['HOSTSIM_JIT_INCLUDE'] = inc
        # the library's kernel cache holds HOST objects here: one directory per version of the stand-in
        os.environ.setdefault('SDFB200_JIT_CACHE', inc.replace('jitinc_', 'jitcache_') + ('_asan' if os.environ.get('HOSTSIM_SANITIZE') else ''))
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'fullsize: GPU parity at the sizes BASELINE.json states (minutes of CPU oracle time; also marked gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
