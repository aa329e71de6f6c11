"""CPU-only: runs the kernel's own per-thread step functions and plans thread-by-thread on the host
(tests/emulate_tile.cu) against the oracle.  This is what lets the CUDA path be debugged without a GPU."""
import os
import subprocess

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_kernel_emulation_matches_oracle(tmp_path):
    ol.load_oracle()
    exe = str(tmp_path / "emulate_tile")
    cmd = ["/usr/local/cuda/bin/nvcc", "-Wno-deprecated-gpu-targets", "-O2", "-o", exe, os.path.join(HERE, "emulate_tile.cu"),
           "-I" + os.path.join(ROOT, "fastecc_b200", "csrc"), "-I" + os.path.join(ROOT, "oracle"),
           os.path.join(ROOT, "oracle", "liboracle.so"), "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")]
    subprocess.run(cmd, check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, cwd=HERE)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "EMULATION OK" in r.stdout
