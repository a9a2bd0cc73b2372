"""Builds tests/cabi_client.c (a plain C99 client of include/sdfb200.h) against the in-tree library."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_c_client(tmp_path):
    if shutil.which('gcc') is None:
        pytest.skip('no gcc')
    exe = os.path.join(str(tmp_path), 'cabi_client')
    lib_dir = os.path.join(ROOT, 'sdf_b200')
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror', '-O1',
                    '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(lib_dir, 'csrc'),
                    os.path.join(ROOT, 'tests', 'cabi_client.c'), '-o', exe,
                    '-L', lib_dir, '-lsdfb200', '-Wl,-rpath,' + lib_dir], check=True)
    return exe
