"""Build tests/simt_emu/_build/libsvae_emu.so: the library's SIMT kernel SOURCES compiled for the host.

Test infrastructure only (see cuda_emu.h).  The .cu files are not modified: this script writes rewritten copies
into _build/ with
  * kernel<<<grid, block, smem, stream>>>(args)   ->  svae_emu::Launcher(grid, block, smem, stream).run(kernel, args)
  * extern __shared__ T name[];                   ->  T* name = (T*)svae_emu::dyn_smem();
  * the two inline-PTX statements (tanh.approx, %globaltimer) -> tanhf / the host clock
and compiles them with g++ against cuda_emu.h.  tc_gemm.cu and tc_bwd.cu (tcgen05 / TMEM / TMA) run on the host model
of those units in tc_emu.h, which replaces the PTX-wrapper section of tc_ptx.cuh.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "spatial-vae_b200", "csrc")
ASAN = os.environ.get("SVAE_EMU_ASAN") == "1"      # out-of-bounds hunting: run python with LD_PRELOAD=libasan.so
BUILD = os.path.join(HERE, "_build_asan" if ASAN else "_build")
OUT = os.path.join(BUILD, "libsvae_emu.so")
SOURCES = ["api.cu", "sgemm.cu", "step_kernels.cu", "option_kernels.cu", "ingest_kernels.cu"]
HEADERS = ["common.cuh", "kernels.cuh", "first_layer.cuh"]
# tensor-core kernel sources and the number of inline red.global.add.v4.f32 statements each contains
TC_SOURCES = {"tc_gemm.cu": 1, "tc_bwd.cu": 1}

LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;]*>)?)<<<(.*?)>>>\(")
DYN_SMEM = re.compile(r"extern __shared__ (?:__align__\(\d+\) )?(\w+) (\w+)\[\];")
TIMER = re.compile(r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\((\w+)\)\);')


RED_V4 = re.compile(r'asm volatile\("red\.global\.add\.v4\.f32 \[%0\], \{%1, %2, %3, %4\};"\s*::"l"\((.*?)\), "f"\((.*?)\), "f"\((.*?)\),\s*"f"\((.*?)\), "f"\((.*?)\) : "memory"\);', re.S)


def rewrite_tc_header(text: str) -> str:
    """tc_ptx.cuh: the PTX-wrapper section is replaced by the host model (tc_emu.h)."""
    a = text.index("// ---- PTX wrappers")
    b = text.index("// ---- descriptors")
    return text[:a] + "}  // namespace\n}  // namespace svae\n#include \"tc_emu.h\"\nnamespace svae {\nnamespace {\n" + text[b:]


def rewrite_tc(text: str, expect_red: int) -> str:
    """tensor-core kernel sources: the inline red.global.add statements become atomics."""
    text, n = RED_V4.subn(lambda m: "{ float* red_p = %s; atomicAdd(red_p, %s); atomicAdd(red_p + 1, %s); "
                                    "atomicAdd(red_p + 2, %s); atomicAdd(red_p + 3, %s); }" % m.groups(), text)
    if n != expect_red:
        raise RuntimeError(f"simt_emu/build.py: expected {expect_red} red.global.add.v4.f32 statement(s), found {n}")
    return text


def rewrite(text: str) -> str:
    text = LAUNCH.sub(lambda m: f"svae_emu::Launcher({m.group(2)}).run({m.group(1)}, ", text)
    text = DYN_SMEM.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)svae_emu::dyn_smem();", text)
    text = text.replace('asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));', "y = tanhf(x);")
    text = TIMER.sub(lambda m: f"{m.group(1)} = svae_emu::globaltimer();", text)
    text = text.replace('#include "../../include/svae_b200.h"',
                        '#include "%s"' % os.path.join(ROOT, "include", "svae_b200.h"))
    for leftover in ("<<<", "asm(", "asm volatile", "extern __shared__"):
        if leftover in text:
            raise RuntimeError(f"simt_emu/build.py: construct not rewritten: {leftover}")
    return text


def newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    inputs = [os.path.join(CSRC, f) for f in SOURCES + HEADERS + list(TC_SOURCES) + ["tc_ptx.cuh"]] + \
             [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "tc_emu.h", "build.py",
                                              "include/cuda.h")] + \
             [os.path.join(ROOT, "include", "svae_b200.h")]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest(inputs):
        return OUT
    for f in HEADERS:
        with open(os.path.join(CSRC, f)) as src, open(os.path.join(BUILD, f), "w") as dst:
            dst.write(rewrite(src.read()))
    units = []
    for f in SOURCES:
        out = os.path.join(BUILD, f.replace(".cu", ".cpp"))
        with open(os.path.join(CSRC, f)) as src, open(out, "w") as dst:
            dst.write(rewrite(src.read()))
        units.append(out)
    units.append(os.path.join(HERE, "cuda_emu.cpp"))
    with open(os.path.join(CSRC, "tc_ptx.cuh")) as src, open(os.path.join(BUILD, "tc_ptx.cuh"), "w") as dst:
        dst.write(rewrite(rewrite_tc_header(src.read())))
    for f, n_red in TC_SOURCES.items():
        out = os.path.join(BUILD, f.replace(".cu", ".cpp"))
        with open(os.path.join(CSRC, f)) as src, open(out, "w") as dst:
            dst.write(rewrite(rewrite_tc(src.read(), n_red)))
        units.append(out)
    flags = ["-std=c++17", "-O2", "-g", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-unknown-pragmas",
             "-I", os.path.join(HERE, "include"), "-I", BUILD, "-I", HERE]
    link = []
    if ASAN:
        flags += ["-fsanitize=address", "-fno-omit-frame-pointer", "-O1"]
        link = ["-fsanitize=address"]
    objs = []
    procs = []
    for u in units:
        o = os.path.join(BUILD, os.path.basename(u) + ".o")
        objs.append(o)
        procs.append((u, subprocess.Popen(["g++", *flags, "-c", u, "-o", o], stderr=subprocess.PIPE, text=True)))
    for u, p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"g++ failed on {u}:\n{err[-4000:]}")
    subprocess.run(["g++", "-shared", *link, "-o", OUT, *objs], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
