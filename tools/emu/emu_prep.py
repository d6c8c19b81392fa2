#!/usr/bin/env python
"""tools/emu: rewrite a .cu file of this repository into host C++ for the emulator (emu_cuda.h).
Only two things change: kernel launches `k<<<grid, block, smem, stream>>>(args);` become
`emu::launch("k", grid, block, MODE, [&] { k(args); });` and the few inline-PTX statements get their C++ meaning.
DEVELOPMENT TOOL ONLY."""
import re
import sys

ASM_RULES = [
    # the MUFU approximations are modelled by correctly rounded values: the Newton sequences that follow are exact for any
    # start within the hardware's error bound, so this checks the logic around them (not the GPU's bits)
    (r'asm\("rcp\.approx\.ftz\.f32 %0, %1;" : "=f"\(r\) : "f"\(b\)\);', 'r = 1.0f / b;'),
    (r'asm\("fma\.rn\.f32 %0, %1, %2, 0f00000000;" : "=f"\(q0\) : "f"\(a\), "f"\(rb\)\);', 'q0 = std::fmaf(a, rb, 0.0f);'),
    (r'asm\("rsqrt\.approx\.ftz\.f32 %0, %1;" : "=f"\(r\) : "f"\(x\)\);', 'r = (float)(1.0 / std::sqrt((double)x));'),
    (r'asm\("mul\.ftz\.f32 %0, %1, %2;" : "=f"\(s\) : "f"\(x\), "f"\(r\)\);', 's = x * r;'),
    (r'asm\("mul\.ftz\.f32 %0, %1, 0f3F000000;" : "=f"\(hh\) : "f"\(r\)\);', 'hh = r * 0.5f;'),
    (r'asm volatile\("ld\.volatile\.global\.v2\.f32 \{%0,%1\}, \[%2\];" : "=f"\(v\.x\), "=f"\(v\.y\) : "l"\(p\) : "memory"\);', 'v = emu::ld_volatile(p);'),
    (r'asm volatile\("st\.volatile\.global\.v2\.f32 \[%0\], \{%1,%2\};" ::"l"\(p\), "f"\(v\.x\), "f"\(v\.y\) : "memory"\);', 'emu::st_volatile(p, v);'),
    (r'asm volatile\("ld\.global\.nc\.v8\.f32 \{%0,%1,%2,%3,%4,%5,%6,%7\}, \[%8\];"\s*:[^;]*?: "l"\(p\)\);', 'c.top = p[0]; c.bot = p[1];'),
    (r'asm volatile\("prefetch\.global\.L1 \[%0\];" ::"l"\(([^;]*)\)\);', r'(void)(\1);'),
    # cp.async helpers: the emulator build takes their S360_EMU branch (synchronous copies); the PTX branch is never compiled
    (r'asm volatile\("cp\.async\.[^;]*?;"[^;]*?\);', ';'),
    (r'asm\("min\.f32 %0, %1, %2, %3;" : "=f"\(r\) : "f"\(a\), "f"\(b\), "f"\(c\)\);[^\n]*', 'r = fminf(fminf(a, b), c);'),
    (r'asm\("max\.f32 %0, %1, %2, %3;" : "=f"\(r\) : "f"\(a\), "f"\(b\), "f"\(c\)\);', 'r = fmaxf(fmaxf(a, b), c);'),
]


def match_close(s, i, open_ch, close_ch):
    depth = 0
    while i < len(s):
        if s[i] == open_ch:
            depth += 1
        elif s[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + open_ch)


def kernel_modes(src):
    modes = {}
    for m in re.finditer(r'__global__\s+void\s*(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(', src):
        name = m.group(1)
        close = match_close(src, m.end() - 1, '(', ')')
        body_start = src.index('{', close)
        body = src[body_start:match_close(src, body_start, '{', '}')]
        if 'ldVolatileF2' in body:     # inter-block hand-off: every warp of the grid resident
            modes[name] = 'emu::GRID_WARPS'
        elif '__syncthreads' in body or '__shfl' in body:   # block at a time; shuffles resolve per warp inside the block

            modes[name] = 'emu::BLOCK_BARRIER'
        else:
            modes[name] = 'emu::SEQUENTIAL'
    return modes


def split_args(s):
    out, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur.strip()); cur = ''
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src, modes):
    out, pos = [], 0
    while True:
        i = src.find('<<<', pos)
        if i < 0:
            out.append(src[pos:]); break
        # callee: identifier with optional template argument list, scanning left from '<<<'
        j = i
        if src[j - 1] == '>':
            depth, j = 0, j - 1
            while True:
                if src[j] == '>': depth += 1
                elif src[j] == '<':
                    depth -= 1
                    if depth == 0: break
                j -= 1
        k = j
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] == '_'):
            k -= 1
        callee = src[k:i]
        base = re.match(r'\w+', callee).group(0)
        cfg_end = src.index('>>>', i)
        cfg = split_args(src[i + 3:cfg_end])
        a0 = src.index('(', cfg_end)
        a1 = match_close(src, a0, '(', ')')
        semi = src.index(';', a1)
        args = src[a0 + 1:a1]
        out.append(src[pos:k])
        out.append(f'emu::launch("{base}", {cfg[0]}, {cfg[1]}, {modes[base]}, [&] {{ {callee}({args}); }});')
        pos = semi + 1
    return ''.join(out)


def main(inp, outp):
    src = open(inp).read()
    # device-only code the emulation cannot express (TMA, mbarrier) is fenced by `#ifndef S360_EMU ... #endif  // S360_EMU`
    src = re.sub(r'#ifndef S360_EMU\n.*?#endif(?:  // S360_EMU)?\n', '', src, flags=re.S)
    for pat, rep in ASM_RULES:
        src, n = re.subn(pat, rep, src, flags=re.S)
    if 'asm' in re.sub(r'//[^\n]*', '', src).replace('emu::', ''):
        left = [l for l in src.splitlines() if re.search(r'\basm\b', l) and not l.strip().startswith('//')]
        if left:
            sys.exit('emu_prep: inline PTX without a rule:\n' + '\n'.join(left))
    modes = kernel_modes(src)
    src = rewrite_launches(src, modes)
    open(outp, 'w').write('// generated by tools/emu/emu_prep.py from ' + inp + '\n#define S360_EMU 1\n' + src)
    print({k: v.split('::')[1] for k, v in modes.items()})


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
