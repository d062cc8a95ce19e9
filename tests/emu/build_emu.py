"""TEST INFRASTRUCTURE ONLY -- build the CPU emulation of liblaplace_b200 for `pytest -m "not gpu"`.

The kernel sources under laplace_gnn_recommendation_b200/csrc/ are taken AS THEY ARE and compiled with g++ against
tests/emu/include/cuda_runtime.h (a warp-lockstep fiber emulation of the CUDA subset they use).  Two purely syntactic
rewrites happen on the way, on copies under tests/emu/_build/gen/:

  * ``kernel<<<grid, block, smem, stream>>>(args);``  ->  ``emu::launch(grid, block, smem, stream, [&]() { kernel(args); });``
  * the bodies of the handful of helper functions written in inline PTX (cache-hinted loads/stores, red.global.add.v4,
    cp.async) are replaced by their plain C++ meaning (table ``ASM_BODIES``); any other ``asm`` makes the build fail.

For csrc/peer.cu the peers are plain host buffers of one process and the NVSwitch multicast address is a key registered
with emu_multicast_bind() (engine.cpp).  The result, tests/emu/_build/liblaplace_b200_emu.so, exports the same C ABI as the real library and is loaded ONLY by
tests (tests/emu/harness.py); the product package never references it.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(REPO, "laplace_gnn_recommendation_b200", "csrc")
HEADER = os.path.join(REPO, "include", "laplace_b200.h")
BUILD = os.path.join(HERE, "_build")
GEN = os.path.join(BUILD, "gen")
# LGB_EMU_ASAN=1: AddressSanitizer build (own output directory) for the memcheck run described in tests/emu/README.md
ASAN = os.environ.get("LGB_EMU_ASAN") == "1"
UBSAN = os.environ.get("LGB_EMU_UBSAN") == "1"      # -fsanitize=undefined build (signed overflow, misaligned vector access, bad shifts)
if ASAN or UBSAN:
    BUILD = os.path.join(HERE, "_build", "asan" if ASAN else "ubsan")
    GEN = os.path.join(BUILD, "gen")
LIB_PATH = os.path.join(BUILD, "liblaplace_b200_emu.so")
SKIP = set()

# helper name -> C++ body with the meaning of the PTX it wraps
ASM_BODIES = {
    "ld_stream_i32": "return *p;",
    "ld_stream_f32": "return *p;",
    "ld_stream_f4": "return *p;",
    "ld_gather_f4": "return *p;",
    "st_f4": "*p = v;",
    "red_add_f4": "p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w;",
    "cp_async_16": "*reinterpret_cast<float4*>(smem_dst) = *reinterpret_cast<const float4*>(gmem_src);",
    "prefetch_l2": "",
    "ld_gather_f8": "f4x2 v; v.a = p[0]; v.b = p[1]; return v;",
    "ld_stream_f8": "f4x2 v; v.a = p[0]; v.b = p[1]; return v;",
    "st_f8": "p[0] = a; p[1] = b;",
    "l2_policy_evict_first": "return 0;",
    "ld_stream_i32_hint": "return *p;",
    "ld_stream_f32_hint": "return *p;",
    "ld_stream_f4_hint": "return *p;",
    "ld_once_f4": "return *p;",
    "ld_once_f4_hint": "return *p;",
    "ld_once_f8": "f4x2 v; v.a = p[0]; v.b = p[1]; return v;",
    "cp_async_commit": "",
    "pin_f4": "",
    "ld_cg_f4": "return *p;",
    "dyn_smem_f4": "static float4 buf[16384]; return buf;",      # 256 KB: blocks run one at a time
    "cp_async_wait": "",
    # NVSwitch multicast (csrc/peer.cu): the "multicast address" is a key registered with emu_multicast_bind(); a load-reduce
    # sums the peers' copies, a store writes all of them
    "multimem_ld_reduce_f4": "return emu::mc_ld_reduce(mc);",
    "multimem_st_f4": "emu::mc_st(mc, v);",
    # signal slots of the in-kernel barrier (only reachable with barriers on, which the single-process emulation never asks for)
    "cas_sys": "uint32_t old = *p; if (old == cmp) *p = val; return old;",
    "fence_sys": "",
    "trap_now": "",
}



def _skip_string(src: str, i: int) -> int:
    """src[i] is a quote: return the index just past the closing quote."""
    q = src[i]
    if q == '"' and src[i - 1] == "R":   # raw string R"( ... )"
        end = src.index(')"', i)
        return end + 2
    i += 1
    while src[i] != q:
        i += 2 if src[i] == "\\" else 1
    return i + 1


def _match(src: str, i: int, open_c: str, close_c: str) -> int:
    """src[i] == open_c: index of the matching close_c (string literals and comments skipped)."""
    depth = 0
    n = len(src)
    while i < n:
        c = src[i]
        if c in "\"'":
            i = _skip_string(src, i)
            continue
        if src.startswith("//", i):
            i = src.index("\n", i)
            continue
        if src.startswith("/*", i):
            i = src.index("*/", i) + 2
            continue
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError(f"unbalanced {open_c}{close_c}")


def _split_top(s: str):
    out, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
    out.append("".join(cur).strip())
    return out


def rewrite_launches(src: str, name: str) -> str:
    out, pos = [], 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier (with ::) and optional balanced template argument list, scanning backwards
        j = k
        while src[j - 1].isspace():
            j -= 1
        if src[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] in "_:"):
            j -= 1
        kernel = src[j:k].strip()
        e = src.index(">>>", k)
        cfg = _split_top(src[k + 3:e])
        if not 2 <= len(cfg) <= 4:
            raise ValueError(f"{name}: launch configuration {cfg!r}")
        cfg += ["0"] * (4 - len(cfg))
        a0 = e + 3
        while src[a0].isspace():
            a0 += 1
        if src[a0] != "(":
            raise ValueError(f"{name}: no argument list after >>> near {src[k - 40:k + 40]!r}")
        a1 = _match(src, a0, "(", ")")
        args = src[a0 + 1:a1]
        out.append(src[pos:j])
        out.append(f"emu::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}, [&]() {{ {kernel}({args}); }})")
        pos = a1 + 1


def rewrite_asm_helpers(src: str, name: str) -> str:
    for fn, body in ASM_BODIES.items():
        for m in list(re.finditer(r"\b" + fn + r"\s*\(", src))[::-1]:
            close = _match(src, m.end() - 1, "(", ")")
            k = close + 1
            while src[k].isspace():
                k += 1
            if src[k] != "{":
                continue            # a call, not the definition
            end = _match(src, k, "{", "}")
            if "asm" not in src[k:end]:
                continue
            src = src[:k] + "{ " + body + " }" + src[end + 1:]
    leftover = re.search(r"\basm\b", re.sub(r"//[^\n]*", "", src))
    if leftover:
        raise ValueError(f"{name}: inline asm outside the known helpers (add it to ASM_BODIES): "
                         f"{src[leftover.start() - 80:leftover.start() + 80]!r}")
    return src


def transform(src: str, name: str) -> str:
    src = src.replace('#include "../../include/laplace_b200.h"', f'#include "{HEADER}"')
    return rewrite_launches(rewrite_asm_helpers(src, name), name)


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in sources()] + [HEADER, os.path.abspath(__file__), os.path.join(HERE, "engine.cpp")]
    for root, _, files in os.walk(os.path.join(HERE, "include")):
        deps += [os.path.join(root, f) for f in files]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps):
        return LIB_PATH
    os.makedirs(GEN, exist_ok=True)
    for f in os.listdir(GEN):
        os.remove(os.path.join(GEN, f))
    units = []
    for f in sources():
        if f in SKIP:
            continue
        text = transform(open(os.path.join(CSRC, f)).read(), f)
        dst = os.path.join(GEN, f if f.endswith(".cuh") else f[:-3] + ".cpp")
        open(dst, "w").write(text)
        if dst.endswith(".cpp"):
            units.append(dst)
    units += [os.path.join(HERE, "engine.cpp")]
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-unknown-pragmas",
             "-Wno-unused-variable", "-I", os.path.join(HERE, "include"), "-I", GEN]
    if ASAN:
        flags += ["-O1", "-fsanitize=address", "-fno-omit-frame-pointer"]
    if UBSAN:
        flags += ["-O1", "-fsanitize=undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"]

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        r = subprocess.run(["g++"] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emulator build failed for {src}:\n" + r.stdout + r.stderr)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, len(units))) as ex:
        objs = list(ex.map(compile_one, units))
    r = subprocess.run(["g++", "-shared", "-o", LIB_PATH + ".tmp"] + (["-fsanitize=address"] if ASAN else []) + (["-fsanitize=undefined"] if UBSAN else []) + objs + ["-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator link failed:\n" + r.stdout + r.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)   # atomic: concurrent test workers never see a half-written library
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
