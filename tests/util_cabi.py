"""Builds tests/cabi_client.c (a plain C99 client of include/sdfb200.h) against the in-tree library."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c_client(tmp_path):
    """Links against the library the Python binding of this process uses (sdf_b200/libsdfb200.so; the host-compiled
    one under pytest --hostsim)."""
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    from sdf_b200 import _abi
    exe = os.path.join(str(tmp_path), 'cabi_client')
    lib_dir, lib_file = os.path.split(_abi.LIB_PATH)
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror', '-O1',
                    '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(ROOT, 'sdf_b200', 'csrc'),
                    os.path.join(ROOT, 'tests', 'cabi_client.c'), '-o', exe,
                    '-L', lib_dir, '-l:' + lib_file, '-Wl,-rpath,' + lib_dir], check=True)
    return exe
