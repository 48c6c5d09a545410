"""Rewrites the CUDA-only syntax of a candmc_b200/csrc/*.cu file so that g++ can compile it for the CPU functional
simulator (tests/cpusim/README.md) — TEST INFRASTRUCTURE, never part of the product.

    kernel<T...><<<grid, block, smem, stream>>>(args...);   ->   CPUSIM_LAUNCH((kernel<T...>), grid, block, smem, stream, args...);
    extern __shared__ double sw[];                          ->   double* sw = (double*)::cpusim::dyn_smem();

Everything else (threadIdx, __syncthreads, __ldg, atomicAdd, __shfl_xor_sync, __shared__, __launch_bounds__) is handled by
macros / functions in sim_device.h, which the build force-includes.  The product sources are not modified.
"""
import re
import sys


def _match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'."""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses after a kernel launch")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


_LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;(){}]*>)?)\s*<<<")
_EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;")


# the few inline-asm statements outside common.cuh's helpers (gemm_f64.cu): rewritten to their meaning
_ASM_RULES = [
    (re.compile(r'asm volatile\("setmaxnreg\.(?:dec|inc)\.sync\.aligned\.u32 %0;"\s*::\s*"n"\([\w:]+\)\);'), "/* setmaxnreg: registers are not modelled */"),
    (re.compile(r'asm volatile\("bar\.sync (\d+), (\d+);"\s*:::\s*"memory"\);'), r"::cpusim::named_barrier(\1, \2);"),
    (re.compile(r'asm volatile\("ld\.acquire\.sys\.global\.u32 %0, \[%1\];"\s*:\s*"=r"\((\w+)\)\s*:\s*"l"\(([^)]+)\)\s*:\s*"memory"\);'),
     r"\1 = __atomic_load_n(reinterpret_cast<const uint32_t*>(\2), __ATOMIC_ACQUIRE);"),
    (re.compile(r'asm volatile\("st\.release\.sys\.global\.u32 \[%0\], %1;"\s*::\s*"l"\((.+?)\),\s*"r"\(([^)]+)\)\s*:\s*"memory"\);'),
     r"__atomic_store_n(reinterpret_cast<uint32_t*>(\1), (uint32_t)(\2), __ATOMIC_RELEASE);"),
    (re.compile(r'asm volatile\("red\.release\.sys\.global\.add\.u32 \[%0\], 1;"\s*::\s*"l"\(([^)]+)\)\s*:\s*"memory"\);'),
     r"__atomic_fetch_add(reinterpret_cast<uint32_t*>(\1), 1u, __ATOMIC_RELEASE);"),
]


def translate(text):
    for rx, rep in _ASM_RULES:
        text = rx.sub(rep, text)
    if "asm volatile" in text or "asm(" in text:
        raise ValueError("inline asm the simulator has no rule for")
    text = _EXTERN_SHARED.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)::cpusim::dyn_smem();", text)
    out, pos = "", 0
    while True:
        m = _LAUNCH.search(text, pos)
        if not m:
            out += text[pos:]
            break
        end_cfg = text.index(">>>", m.end())
        cfg = _split_top(text[m.end():end_cfg])
        while len(cfg) < 4:
            cfg.append("0")
        k = end_cfg + 3
        while text[k].isspace():
            k += 1
        assert text[k] == "(", f"kernel launch without argument list near: {text[m.start():m.start() + 80]!r}"
        close = _match_paren(text, k)
        args = text[k + 1:close].strip()
        out += text[pos:m.start()]
        out += f"CPUSIM_LAUNCH(({m.group(1)}), {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}{', ' + args if args else ''})"
        pos = close + 1
    return out


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        body = translate(f.read())
    with open(dst, "w") as f:
        f.write(f'#line 1 "{src}"\n' + body)
