"""A selection of the `-m gpu` tests, UNCHANGED, run where there is no GPU: against the whole C-ABI library compiled
for the host (tests/util_hostsim.py build_library: sdfb200.cu's own orchestration -- work-list sizing, rounds,
retries, counters, host / device output paths -- over the kernels on the SIMT stand-in and a stand-in CUDA runtime).

The GPU tests run in a child pytest process started with ``--hostsim`` (tests/conftest.py points the ctypes binding
at the host-compiled library before it is first loaded), so no handle ever crosses from one library to the other.
What runs here also runs, identically, on the B200 in the `-m gpu` tier; what cannot run here: anything that needs
torch.cuda and the sizes that only make sense on a GPU.  Run-time specialised kernels DO run (the source the library
generates is compiled for the host by a stand-in NVRTC) and so does the one-graph step (a captured stream records
closures, a graph launch replays them).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SELECTIONS = {
    'engine': (['tests/test_gpu_parity.py'],
               'interp and (edge_cases or stl_bytes or user_closure or sparse_off or (generate_matches_oracle and sphere) '
               'or (exact_mode_equals and csg) or skip_flags)'),
    'closures_textures_c_client': (['tests/test_gpu_zzz_texture.py', 'tests/test_gpu_closure.py', 'tests/test_gpu_zzz_cabi_client.py'], ''),
    'weld_windows': (['tests/test_gpu_weld.py', 'tests/test_gpu_windows.py'], 'sphere or csg'),
    'guardband': (['tests/test_gpu_guardband.py'], 'twist_fast or scaled_up or square_jump'),
    # the library's run-time specialisation end to end: jit_source(), the forced thread-run shapes, background compiles,
    # the kernel cache, cudaLaunchKernel's argument arrays -- its generated source compiled for the host by a stand-in
    # NVRTC (tests/hostsim/fake_nvrtc.cpp)
    'specialised_kernels': (['tests/test_gpu_parity.py', 'tests/test_gpu_grouping.py'],
                            '(jit and generate_matches_oracle and exact and (sphere or csg)) or (sample_grouping and cyl_x)'),
    # paths of sdfb200.cu's orchestration that only sizes beyond any emulation reach on a GPU (tests/hostsim_paths.py)
    'host_cut_rounds_overflow_retry_pipelined_output_slabs': (['tests/hostsim_paths.py'], ''),
}


@pytest.mark.parametrize('name', list(SELECTIONS))
def test_gpu_tests_pass_on_the_host_compiled_library(name):
    files, expr = SELECTIONS[name]
    cmd = [sys.executable, '-m', 'pytest', '-m', 'gpu', '--hostsim', '-q', '-x', '-p', 'no:cacheprovider'] + files
    if expr:
        cmd += ['-k', expr]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert ' passed' in r.stdout and ' failed' not in r.stdout, tail
