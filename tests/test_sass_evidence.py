"""CPU: the built library really contains what DESIGN.md says the kernels are made of.  cuobjdump -sass of
fastecc_b200/libfastecc_b200.so (sm_100a cubins only): the headline pass kernels take their quotient from the FP64 unit
(I2F.F64.U32 + DFMA.RM + VIADDMNMX.U32, one each per product, and no IMAD.HI in the butterflies), load their tiles and tables
with TMA (UTMALDG / UBLKCP + SYNCS mbarrier instructions) and do not spill."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastecc_b200", "libfastecc_b200.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("needs cuobjdump and the built library")
    elf = subprocess.run([exe, "-lelf", LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); kernels[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            kernels[name].append(line)
    return kernels


def _kernel(sass, pattern):
    hits = [k for k in sass if pattern in k]
    assert len(hits) == 1, hits
    return sass[hits[0]]


@pytest.mark.parametrize("pattern", ["ntt_pass_kernelILi10ELi2ELi1E", "ntt_pass_kernelILi9ELi1ELi1E", "ntt_pass_kernelILi10ELi2ELi2E", "ntt_pass_kernelILi9ELi1ELi2E"])
def test_headline_kernels_use_the_fp64_quotient_and_tma(sass, pattern):
    body = _kernel(sass, pattern)
    text = "\n".join(body)
    n_cvt, n_fma, n_fix = text.count("I2F.F64.U32"), text.count("DFMA.RM"), text.count("VIADDMNMX.U32")
    assert n_cvt >= 300 and n_cvt == n_fma and n_fix >= n_cvt                  # one conversion, one DFMA, one add-min per product
    assert text.count("IMAD.HI") <= 16                                            # address arithmetic only: none in the butterflies
    assert "UTMALDG" in text and "UBLKCP" in text and "SYNCS" in text            # TMA tensor + bulk loads, mbarrier
    assert sum(1 for l in body if "LDL" in l or "STL" in l) <= 4                  # no spilling to speak of


def test_library_contains_every_kernel_family(sass):
    names = " ".join(sass)
    for k in ("ntt_pass_kernel", "build_tables_kernel", "small_dft_kernel", "radix_pass_kernel", "row_scale_kernel", "shard_barrier_kernel",
              "leaves_kernel", "pairmul_kernel", "gather_scale_kernel"):
        assert k in names, k
