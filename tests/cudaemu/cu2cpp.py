#!/usr/bin/env python
"""cu2cpp.py -- TEST INFRASTRUCTURE (see cuda_emu.h).  Rewrites the two pieces of CUDA syntax
g++ cannot parse so that a kernel source compiles for the host emulator unchanged otherwise:

    kernel<<<grid, block[, smem[, stream]]>>>(args);
        -> emu::launch("kernel", dim3(grid), dim3(block), smem, [&]() { kernel(args); });
    extern __shared__ [__align__(N)] T name[];
        -> T *name = reinterpret_cast<T *>(emu::dyn_smem());

usage: cu2cpp.py in.cu out.cpp
"""
import re
import sys


def split_top(s):
    """split on commas that are not inside (), [], {} or <>-free contexts"""
    out, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    out.append("".join(cur).strip())
    return out


def kernel_expr_start(src, end):
    """index where the kernel expression that ends at `end` (exclusive) starts:
    identifier[::identifier...][<template args>]"""
    i = end
    while i > 0 and src[i - 1].isspace():
        i -= 1
    if src[i - 1] == ">":
        depth = 0
        while i > 0:
            i -= 1
            if src[i] == ">":
                depth += 1
            elif src[i] == "<":
                depth -= 1
                if depth == 0:
                    break
    while i > 0 and (src[i - 1].isalnum() or src[i - 1] in "_:"):
        i -= 1
    return i


def transform(src):
    out = []
    pos = 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            break
        e = src.find(">>>", k)
        assert e > 0, "unterminated launch configuration"
        s = kernel_expr_start(src, k)
        kern = src[s:k].strip()
        cfg = split_top(src[k + 3:e])
        assert 2 <= len(cfg) <= 4, cfg
        smem = cfg[2] if len(cfg) > 2 else "0"
        # argument list
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", src[a:a + 20]
        depth, b = 0, a
        while True:
            if src[b] == "(":
                depth += 1
            elif src[b] == ")":
                depth -= 1
                if depth == 0:
                    break
            b += 1
        args = src[a + 1:b]
        out.append(src[pos:s])
        name = kern.replace('"', "")
        out.append(f'emu::launch("{name}", dim3({cfg[0]}), dim3({cfg[1]}), (size_t)({smem}), [&]() {{ {kern}({args}); }})')
        pos = b + 1
    text = "".join(out)
    text = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
                  r"\1 *\2 = reinterpret_cast<\1 *>(emu::dyn_smem());", text)
    return text


if __name__ == "__main__":
    src = open(sys.argv[1]).read()
    open(sys.argv[2], "w").write(f'#line 1 "{sys.argv[1]}"\n' + transform(src))
