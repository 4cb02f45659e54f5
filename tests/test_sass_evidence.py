"""The built library contains the instructions the design relies on (no GPU needed: cuobjdump on librt3cuda.so).

The sweep's speed hinges on two code-generation facts that a harmless-looking source change can undo without any
functional symptom: the level-1 test must be packed FMAs (SASS FFMA2) whose primitive operand
comes through the uniform datapath (LDCU.64 loads from the constant bank; ptxas falls back to per-thread LDC as soon as
the control flow around the sweep stops looking convergent to it, which costs 20 %), and large scenes must be staged by
bulk asynchronous copies (UBLKCP)."""
import re
import shutil
import subprocess

import pytest

from rt3_b200 import abi

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def sass(built):
    text = subprocess.run(["cuobjdump", "-sass", abi.CORE_LIB_PATH], check=True, capture_output=True, text=True).stdout
    kernels = {}
    for block in text.split("Function : ")[1:]:
        name, _, body = block.partition("\n")
        kernels[name.strip()] = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", body, flags=re.M)
    return kernels


def kernel(sass, *needles):
    names = [n for n in sass if all(x in n for x in needles)]
    assert len(names) == 1, names
    return sass[names[0]]


def test_built_for_sm_100a(built):
    out = subprocess.run(["cuobjdump", "-lelf", abi.CORE_LIB_PATH], check=True, capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("needles", [("pathtrace_kernelILb1ELb1ELb0ELi0ELb0",), ("pathtrace_kernelILb1ELb1ELb0ELi0ELb1",), ("pathtrace_kernelILb1ELb0ELb0",),
                                     ("reference_kernelILb1ELb1ELb0",), ("reference_kernelILb1ELb0ELb0",)])
def test_constant_bank_sweep_uses_packed_fma_with_uniform_operands(sass, needles):
    ops = kernel(sass, *needles)
    ffma2 = sum(op.startswith("FFMA2") for op in ops)
    ldcu64 = sum(op == "LDCU.64" for op in ops)
    ldc64 = sum(op == "LDC.64" for op in ops)
    assert ffma2 >= 96, f"{ffma2} FFMA2"                      # 16 pairs x 2 rays x 3 in the unrolled word (+ the partial-word loop)
    assert ldcu64 >= 48, f"only {ldcu64} LDCU.64 ({ldc64} LDC.64): the records are no longer read through uniform registers"
    # kernel parameters (camera, pointers) are read with per-thread LDC.64 outside the sweep: a dozen in the path tracer plus as many in
    # its out-of-line tail routine (sweep_slots_by_primitive), a few more in the reference-mode kernel; a sweep that fell back to LDC
    # shows 60 or more (and hardly any LDCU.64, which the assertion above catches first)
    assert ldc64 < (60 if "ELb1" in needles[0][-4:] else 40), f"{ldc64} LDC.64: per-thread constant loads in the sweep"  # the BEAM kernel reads the camera once more
    assert sum(op.startswith("SHF.L.W") for op in ops) >= 64  # sign bits into the survivor masks


@pytest.mark.parametrize("needles", [("pathtrace_kernelILb0ELb1ELb0",), ("reference_kernelILb0ELb0ELb0",)])
def test_streamed_sweep_stages_tiles_with_bulk_copies(sass, needles):
    ops = kernel(sass, *needles)
    assert sum(op.startswith("UBLKCP") for op in ops) >= 2, "no cp.async.bulk (UBLKCP) in the streamed kernel"
    assert sum(op.startswith("FFMA2") for op in ops) >= 96
    assert any(op.startswith("SYNCS") for op in ops), "no mbarrier instructions"


@pytest.mark.parametrize("needle", ["pathtrace_kernelILb1ELb1ELb0ELi0ELb0", "pathtrace_kernelILb1ELb1ELb0ELi0ELb1"])
def test_accumulation_is_a_64_bit_reduction(sass, needle):
    ops = kernel(sass, needle)
    assert any(op.startswith("RED") and "64" in op for op in ops)


@pytest.fixture(scope="module")
def resources(built):
    """{mangled kernel name: (registers, spill-free?)} from cuobjdump -res-usage."""
    text = subprocess.run(["cuobjdump", "-res-usage", abi.CORE_LIB_PATH], check=True, capture_output=True, text=True).stdout
    out = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", text):
        out[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(5)))
    return out


# (instantiation, register cap that keeps the resident CTAs the launch bounds ask for, largest stack frame: the traversal's 128 x 8-byte
# stack + the path state, or the tail routine's frame in the sweep kernels)
@pytest.mark.parametrize("needle,max_regs,max_stack", [
    ("pathtrace_kernelILb1ELb1ELb0ELi0ELb0", 73, 256),    # sweep, resident spheres: 7 CTAs of 128 threads per SM
    ("pathtrace_kernelILb1ELb1ELb0ELi0ELb1", 73, 256),    # ... with the primary rays through candidate lists (the headline kernel)
    ("pathtrace_kernelILb1ELb0ELb1ELi0ELb0", 85, 1200),   # hierarchy, 6 CTAs per SM
    ("pathtrace_kernelILb1ELb0ELb1ELi3ELb0", 85, 1200),   # ... sorted traversal
    ("pathtrace_kernelILb1ELb0ELb1ELi0ELb1", 85, 1200),   # ... with the beams (sphere scenes, small meshes)
    ("pathtrace_kernelILb1ELb0ELb1ELi3ELb1", 85, 1200),   # ... sorted traversal with the beams (RT3_BEAM_BVH=1)
])
def test_render_kernels_keep_their_occupancy(resources, needle, max_regs, max_stack):
    names = [n for n in resources if needle in n]
    assert len(names) == 1, names
    regs, stack, local = resources[names[0]]
    assert regs <= max_regs, f"{regs} registers: fewer resident CTAs per SM than the kernel was tuned for"
    assert stack <= max_stack and local == 0, f"stack {stack} B, local {local} B: the kernel spills"
