"""TEST INFRASTRUCTURE ONLY: compile the kernel sources as plain C++ (-DADM_HOST_EMU) into tests/emu/libadm_emu.so
so that the kernel LOGIC can be unit-tested on the CPU (tests/test_kernel_logic_emu.py).  Never used by the product."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'adorym_b200', 'csrc')
LIB = os.path.join(HERE, 'libadm_emu.so')


def build():
    srcs = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    deps = srcs + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(HERE, 'emu_runtime.h')]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ['g++', '-std=c++20', '-O2', '-fPIC', '-shared', '-pthread', '-DADM_HOST_EMU', '-I', HERE, '-I', CSRC,
           '-Wno-unknown-pragmas', '-o', LIB]
    for s in srcs:
        cmd += ['-x', 'c++', s]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('emu build failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build())
