"""CPU: the hot kernels of the bench step exist in libserac_b200.so as sm_100a code with the resource usage their
occupancy is planned on (cuobjdump -res-usage; no device needed): a spill or a register jump after a refactor shows up
here instead of as a silent slowdown on the B200.  Numbers: profiles/r4_summary.md, DESIGN.md section 4."""
import re
import shutil
import subprocess

import pytest

from serac_b200 import capi

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump (CUDA toolkit)")

# mangled-name fragment -> (max registers, what the budget is for)
HOT = {
    # config A Jacobian action, symmetric tangent storage: 5 CTAs x 64 threads per SM need <= 204 registers
    "jvp_stream_kernelILi3ELi3ELi3ELi3ELi3ELi3ELi24ELi2ELi5E": (176, "5 resident CTAs per SM"),
    # config B (p = 3) Jacobian action: 3 CTAs x 128 threads
    "jvp_stream_kernelILi3ELi4ELi4ELi3ELi4ELi3ELi24ELi2ELi3E": (170, "3 resident CTAs per SM"),
    # config A evaluation with derivative store (NeoHookean, geometric nonlinearity on, tangent fast path): 16 CTAs per SM
    ("evaluation_kernelINS0_10SignatureTILNS_8GeometryE5ELi3ELi0ENS_2H1ILi2ELi3EEEJNS4_ILi1ELi3EEES5_S5_EEELi1E"
     "NS_35ShapeAwareIntegrandWrapperWithStateINS_21MaterialStressFunctorINS_15solid_mechanics10NeoHookeanELNS_8FlagModeE2EEE"
     "Li3ES5_JS5_S5_EEENS_5EmptyELi1ELi16ELb1E"): (128, "16 resident CTAs per SM"),
    # config A fused assembly: 12 warps per SM
    "assemble_rows_kernelILi3ELi3ELi3ELi3ELi3ELi24ELi2E": (170, "12 resident warps per SM"),
}


@pytest.fixture(scope="module")
def res_usage():
    out = subprocess.run(["cuobjdump", "-res-usage", capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    usage, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
            name = None
    return usage


def test_library_holds_sm_100a_code_only():
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("fragment", list(HOT))
def test_hot_kernel_registers_and_spills(res_usage, fragment):
    max_regs, why = HOT[fragment]
    hits = {k: v for k, v in res_usage.items() if fragment in k}
    assert hits, f"kernel not found in the library: {fragment[:60]}..."
    for name, (regs, stack) in hits.items():
        assert stack == 0, f"{name[:80]}: {stack} bytes of stack (register spills)"
        assert regs <= max_regs, f"{name[:80]}: {regs} registers, budget {max_regs} ({why})"
