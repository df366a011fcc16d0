"""CPU test (needs only cuobjdump): the MMA issue loops of the conv kernels must stay in UNIFORM registers.

Round 2 found the conv kernels held at 0.73 of the tensor peak by a `__noinline__` slow path of `mbar_wait`: a possible call
inside the MMA issue loop made the compiler keep the loop's state in thread registers and rebuild the uniform registers
the `UTCHMMA` operands need after it (six `R2UR` per tap).  With the wait fully inline there is no `R2UR` and no `CALL`
between the first and the last `UTCHMMA` of any conv kernel.  This guards the property in the built
library: a regression here costs 8-10 % of the headline number without failing any numerical test."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-enhance-keras_b200", "lib", "libsr100.so")


@pytest.fixture(scope="module")
def sass_functions():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    txt = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = f.split("\n", 1)
        funcs[name.strip()] = body.split("\n")
    return funcs


def _mma_span(lines):
    at = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
    return lines[at[0]:at[-1] + 1] if at else []


def test_conv_mma_issue_loops_have_no_call_and_no_r2ur(sass_functions):
    checked = 0
    for name, lines in sass_functions.items():
        if "conv_tc_pair_kernel" not in name and "conv_tc_kernel" not in name and "conv_tc_chain_kernel" not in name:
            continue
        span = _mma_span(lines)
        assert span, "no UTCHMMA in " + name
        calls = sum("CALL" in l for l in span)
        r2ur = sum("R2UR" in l for l in span)
        assert calls == 0, "a call inside the MMA issue loop of %s" % name
        # the sub-pixel shuffle variants (EPI = 4) interleave a little per-thread epilogue set-up: allow their handful
        limit = 8 if re.search(r"ELi4ELb0ELb0E", name) else 0
        assert r2ur <= limit, "%d R2UR inside the MMA issue loop of %s" % (r2ur, name)
        checked += 1
    assert checked >= 25       # every conv_tc_kernel / conv_tc_pair_kernel / chain instantiation of the library


def test_wgrad_mma_issue_loop_has_no_call_and_no_r2ur(sass_functions):
    """Same property for the filter-gradient kernel: an `if (has_work)` around its issue loop made ptxas treat the loop as
    possibly divergent (102 R2UR in the MMA span, 64 registers); without the branch the loop is uniform (0 R2UR, 45
    registers) and the kernel is 3-11 % faster, bit-identical (profiles/r02_ab_wgrad_uniform_issue_loop.json)."""
    names = [n for n in sass_functions if "wgrad_tc_kernel" in n]
    assert names
    for name in names:
        span = _mma_span(sass_functions[name])
        assert span, "no UTCHMMA in " + name
        assert sum("CALL" in l for l in span) == 0, name
        assert sum("R2UR" in l for l in span) == 0, "%d R2UR inside the MMA issue loop of %s" % (
            sum("R2UR" in l for l in span), name)


def test_waits_are_inline(sass_functions):
    """mbar_wait is __forceinline__ and nothing in ptx.cuh is __noinline__; no tensor-core kernel references vprintf (the old
    slow path printed before trapping)."""
    src = open(os.path.join(ROOT, "image-enhance-keras_b200", "csrc", "ptx.cuh")).read()
    code = "\n".join(l for l in src.split("\n") if not l.lstrip().startswith("//"))
    assert "__noinline__" not in code
    assert re.search(r"__device__ __forceinline__ void mbar_wait\(", code)
    for name, lines in sass_functions.items():
        if "wgrad_tc_kernel" in name or "conv_tc_pair_kernel" in name or "conv_tc_kernel" in name:
            assert not any("vprintf" in l for l in lines), name
