#!/usr/bin/env python3
"""Build-time view of member functions of a reference translation unit that cannot be compiled whole (test infrastructure).

src/backMapping.cpp needs roscpp, tf, GTSAM, small_gicp and PCL's I/O to compile as a file; the scan-to-map member functions this
repository restates need only PCL points, a kd-tree, cv::Mat and Eigen. This script finds each named member function
in the reference file where it lies (by its signature, then by matching braces — no line numbers), and writes their text, unmodified and
in file order, to an include file under oracle/_ref/ (git-ignored build output). oracle/ref_bm_driver.cpp includes that file inside a
class that declares the members the functions use. Nothing from the reference is stored in the repository.

    extract_ref_members.py <reference .cpp> <output .inc> <signature> [<signature> ...]
"""
import re
import sys


def strip_noise(line: str) -> str:
    line = re.sub(r'"(\\.|[^"\\])*"', '""', line)
    line = re.sub(r"'(\\.|[^'\\])'", "''", line)
    return line.split("//", 1)[0]


def find_function(lines, signature):
    hits = [i for i, l in enumerate(lines) if signature in strip_noise(l) and not strip_noise(l).rstrip().endswith(";")]
    if len(hits) != 1:
        raise SystemExit(f"extract_ref_members: expected exactly one definition of `{signature}`, found {len(hits)}")
    start = hits[0]
    depth, seen_open = 0, False
    for i in range(start, len(lines)):
        code = strip_noise(lines[i])
        for ch in code:
            if ch == "{":
                depth += 1; seen_open = True
            elif ch == "}":
                depth -= 1
        if seen_open and depth == 0:
            return start, i
    raise SystemExit(f"extract_ref_members: unbalanced braces after `{signature}`")


def main():
    src, out, sigs = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = open(src, encoding="utf-8", errors="replace").read().split("\n")
    spans = sorted(find_function(lines, s) + (s,) for s in sigs)
    with open(out, "w", encoding="utf-8") as f:
        f.write(f"// GENERATED at build time from {src} — do not commit. Member functions, verbatim, in file order.\n")
        for a, b, s in spans:
            f.write(f"// ---- {s}: lines {a + 1}-{b + 1}\n#line {a + 1} \"{src}\"\n")
            f.write("\n".join(lines[a:b + 1]) + "\n")
    print(f"extract_ref_members: {len(spans)} functions -> {out}")


if __name__ == "__main__":
    main()
