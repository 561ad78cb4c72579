#!/usr/bin/env python
"""TEST TOOLING — NOT PRODUCT CODE. Builds tools/cuda_emu/_build/libkhronos_b200_emu.so: the product's CUDA sources
(khronos_b200/csrc, unmodified apart from a mechanical rewrite of the `kernel<<<...>>>(...)` launch syntax and of
`extern __shared__` declarations) compiled with g++ against the CUDA-on-CPU shim in tools/cuda_emu/include.
Used by tools/cuda_emu/run_parity.py to check kernel *logic* against the oracle where no GPU is available."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "khronos_b200", "csrc")
# build artefacts (rewritten copies of the product sources) live outside the source tree
BUILD_ROOT = os.environ.get("KB_EMU_BUILD_ROOT", "/tmp/khronos_b200_cuda_emu")
OUT = os.path.join(BUILD_ROOT, "_build")
LIB = os.path.join(OUT, "libkhronos_b200_emu.so")

LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)<<<(.+?)>>>\((.*?)\);")
SMEM = re.compile(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];")


def rewrite(src: str) -> str:
    src = LAUNCH.sub(lambda m: f"emu::launch_([=]() {{ {m.group(1)}({m.group(3)}); }}, {m.group(2)});", src)
    src = SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = static_cast<{m.group(1)}*>(emu::dyn_smem());", src)
    return src


def build(verbose=False, asan=None):
    """asan (default: environment KB_EMU_ASAN=1): AddressSanitizer build in _build_asan/ — every "device" allocation is a malloc,
    so out-of-bounds kernel accesses are caught like compute-sanitizer memcheck would catch them. Run with
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0."""
    global OUT, LIB
    if asan is None:
        asan = os.environ.get("KB_EMU_ASAN") == "1"
    ubsan = not asan and os.environ.get("KB_EMU_UBSAN") == "1"   # UndefinedBehaviorSanitizer build in _build_ubsan/ (reports to stderr;
    if asan:                                                        # run with LD_PRELOAD=$(g++ -print-file-name=libubsan.so))
        OUT = os.path.join(BUILD_ROOT, "_build_asan")
        LIB = os.path.join(OUT, "libkhronos_b200_emu.so")
    elif ubsan:
        OUT = os.path.join(BUILD_ROOT, "_build_ubsan")
        LIB = os.path.join(OUT, "libkhronos_b200_emu.so")
    deps = [os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.endswith((".cu", ".cuh", ".h", ".cpp"))]
    deps += [os.path.join(HERE, "emu_runtime.cpp"), os.path.join(HERE, "include", "cuda_runtime.h"), os.path.abspath(__file__),
             os.path.join(ROOT, "include", "khronos_b200.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(os.path.join(OUT, "include"), exist_ok=True)
    srcs = []
    for name in sorted(os.listdir(CSRC)):
        if not name.endswith((".cu", ".cuh", ".h", ".cpp")):
            continue
        text = rewrite(open(os.path.join(CSRC, name)).read())
        text = text.replace('"../../include/khronos_b200.h"', f'"{os.path.join(ROOT, "include", "khronos_b200.h")}"')
        out = os.path.join(OUT, name.replace(".cu", ".cpp") if name.endswith(".cu") else name)
        if name.endswith(".cuh"):
            out = os.path.join(OUT, name)
        open(out, "w").write(text)
        if out.endswith(".cpp"):
            srcs.append(out)
    cmd = ["g++", "-O1" if asan else "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas"]
    if asan:
        cmd += ["-fsanitize=address", "-fno-omit-frame-pointer"]
    elif ubsan:
        cmd += ["-fsanitize=undefined", "-fno-omit-frame-pointer"]
    cmd += [
           "-I", os.path.join(HERE, "include"), "-I", OUT, "-o", LIB, os.path.join(HERE, "emu_runtime.cpp")] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
