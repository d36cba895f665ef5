"""Rewrite the CUDA sources of wolkenbase_b200/csrc into host C++ for the SIMT emulator: every
`kernel<<<grid, block, smem, stream>>>(args);` becomes `simt_launch(grid, block, [&]{ kernel(args); });`.
Everything else is compiled as it stands against tests/simt/shim/cuda_runtime.h.  TEST INFRASTRUCTURE ONLY: the
output (tests/simt/gen/) is a build product, never committed, never on the product path."""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "wolkenbase_b200", "csrc")
GEN = os.path.join(HERE, "gen")


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite(text):
    out, pos, count = [], 0, 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            out.append(text[pos:])
            break
        # kernel name (with optional template arguments) right before <<<
        m = re.search(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*$", text[pos:k])
        assert m, text[k - 60:k + 20]
        name_start = pos + m.start(1)
        e = text.index(">>>", k)
        cfg = split_top(text[k + 3:e])
        j = e + 3
        while text[j].isspace():
            j += 1
        assert text[j] == "(", text[e:e + 40]
        depth, a0 = 0, j
        while True:
            if text[j] == "(":
                depth += 1
            elif text[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        args = text[a0 + 1:j]
        out.append(text[pos:name_start])
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append("simt_launch(%s,%s,%s,[&]{ %s(%s); })" % (cfg[0], cfg[1], smem, m.group(1), args))
        pos = j + 1
        count += 1
    return "".join(out), count


def main():
    os.makedirs(GEN, exist_ok=True)
    total = 0
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh", ".h")):
            continue
        text = open(os.path.join(CSRC, f)).read()
        text, n = rewrite(text)
        total += n
        # dynamic shared memory: `extern __shared__ T name[];` -> a pointer into the launch's buffer
        text = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1 *\2=(\1 *)simt_dynamic_smem();", text)
        text = text.replace('#include "../../include/wolken_b200.h"', '#include "../../../include/wolken_b200.h"')
        name = f + ".cpp" if f.endswith(".cu") else f
        open(os.path.join(GEN, name), "w").write(text)
    print("rewrote %d launches" % total)


if __name__ == "__main__":
    main()
