"""Static checks on the built library's SASS (no GPU needed): the hot kernels really are tcgen05 / bulk-copy kernels, and
their MMA warps issue from converged code (no lane-by-lane waterfall loop around the tensor-core instructions)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from tvqaplus_b200 import _lib


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    out = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    per = collections.defaultdict(collections.Counter)
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name is None:
            continue
        if re.search(r"\bUTC[A-Z]*MMA\b", line):
            per[name]["mma"] += 1
        if "BRA.U.ANY" in line:
            per[name]["waterfall"] += 1
        if "UBLKCP" in line:
            per[name]["bulk"] += 1
        if re.search(r"\bLDTM\b|\bSTTM\b", line):
            per[name]["tmem_ldst"] += 1
        if "FFMA2" in line:
            per[name]["ffma2"] += 1
    return per


def kernels(per, key):
    return {k: v for k, v in per.items() if key in k}


def test_hot_kernels_use_tcgen05_and_bulk_copies(sass):
    for key in ("linear_kernel", "encoder_kernel", "cqa_kernel", "mha_kernel"):
        ks = kernels(sass, key)
        assert ks, "no %s in the library" % key
        for name, c in ks.items():
            assert c["mma"] > 0, "%s has no tcgen05.mma" % name
            assert c["bulk"] > 0, "%s has no cp.async.bulk" % name
            assert c["tmem_ldst"] > 0, "%s never touches tensor memory from registers" % name


def test_mma_warps_issue_from_converged_code(sass):
    """Under a `lane == 0` branch the compiler wraps every tcgen05 instruction in a waterfall loop (BRA.U.ANY): the
    linear / encoder / CQA / MHA kernels must have (almost) none, certainly fewer than one per MMA."""
    for key in ("linear_kernel", "encoder_kernel", "cqa_kernel", "mha_kernel"):
        for name, c in kernels(sass, key).items():
            assert c["waterfall"] <= 4, "%s: %d waterfall loops for %d MMAs" % (name, c["waterfall"], c["mma"])


def test_packed_fp32_arithmetic_is_used(sass):
    for key in ("encoder_kernel", "cqa_kernel", "linear_kernel"):
        for name, c in kernels(sass, key).items():
            assert c["ffma2"] > 0, "%s has no FFMA2" % name
