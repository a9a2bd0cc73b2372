"""Line coverage of the DEVICE sources under the host execution (tests/hostsim/): builds the harness with gcov
instrumentation, runs tests/test_hostsim_cpu.py on it (interpreter evaluator; the specialised-evaluator and sanitizer
tests use other builds) and prints gcov's summary for sdf_kernels.cuh / sdf_ops.cuh.  No GPU.  ~2 min.

    python tools/hostsim_coverage.py [workdir]        .gcov listings are left in the work directory (default /tmp/hostsim_cov)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]


def main():
    import util_hostsim as U
    work = sys.argv[1] if len(sys.argv) > 1 else '/tmp/hostsim_cov'
    src = os.path.join(work, 'src')
    os.makedirs(src, exist_ok=True)
    for f in os.listdir(work):
        if f.endswith(('.gcda', '.gcno', '.gcov', '.so')):
            os.remove(os.path.join(work, f))
    U.patched_sources(src)
    so = os.path.join(work, 'hostsim_cov.so')
    subprocess.run(['g++', '-std=c++17', '-O0', '-g', '--coverage', '-fPIC', '-shared', '-ffp-contract=off', '-fno-strict-aliasing',
                    '-Wno-unknown-pragmas', '-I', src, '-I', U.HS, os.path.join(U.HS, 'harness.cpp'), '-o', so], check=True, cwd=work)
    runner = ("import sys; sys.path[:0] = [%r, %r]; import util_hostsim as U, pytest; "
              "U.build = lambda spec_text=None, sanitize=False: %r; "
              "sys.exit(pytest.main([%r, '-q', '-k', 'not specialised and not sanitizers', '-p', 'no:cacheprovider']))"
              % (ROOT, os.path.join(ROOT, 'tests'), so, os.path.join(ROOT, 'tests', 'test_hostsim_cpu.py')))
    subprocess.run([sys.executable, '-c', runner], check=True, cwd=work)
    out = subprocess.run(['gcov', '-o', work, 'hostsim_cov.so-harness.gcno', '-r', '-s', src], capture_output=True, text=True, cwd=work).stdout
    lines = out.splitlines()
    for i, ln in enumerate(lines):
        if ln.startswith("File '") and ('sdf_kernels.cuh' in ln or 'sdf_ops.cuh' in ln):
            print(ln, '--', lines[i + 1])
    missed = [ln for ln in open(os.path.join(work, 'sdf_kernels.cuh.gcov')) if ln.lstrip().startswith('#####')]
    print('sdf_kernels.cuh: %d lines (over all template instantiations) never executed; listing in %s' % (len(missed), work))


if __name__ == '__main__':
    main()
