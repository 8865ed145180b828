#!/usr/bin/env python3
"""Build-time view of member functions of a reference translation unit that cannot be compiled whole (test infrastructure).

src/backMapping.cpp needs roscpp, tf, GTSAM, small_gicp and PCL's I/O to compile as a file; the scan-to-map member functions this
repository restates need only PCL points, a kd-tree, cv::Mat and Eigen. This script finds each named member function
in the reference file where it lies (by its signature, then by matching braces — no line numbers), and writes their text, unmodified and
in file order, spliced into the driver source in place of its marker line, to STDOUT — the Makefile pipes that straight into the compiler
(`g++ -x c++ -`), so the reference text exists only in the pipe: nothing from the reference is stored in the repository or left on disk.
oracle/ref_bm_driver.cpp is the driver: a class that declares the members the functions use, with the marker where they go.

    extract_ref_members.py <reference .cpp> <driver .cpp> <marker text> <signature> [<signature> ...]   > translation unit
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
    src, driver, marker, sigs = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
    lines = open(src, encoding="utf-8", errors="replace").read().split("\n")
    spans = sorted(find_function(lines, s) + (s,) for s in sigs)
    body = []
    for a, b, s in spans:
        body.append(f"// ---- {s}: lines {a + 1}-{b + 1}\n#line {a + 1} \"{src}\"")
        body.append("\n".join(lines[a:b + 1]))
    out = []
    hits = 0
    for n, l in enumerate(open(driver, encoding="utf-8").read().split("\n")):
        if marker in l and not l.lstrip().startswith("//"):
            hits += 1
            out.extend(body)
            out.append(f"#line {n + 2} \"{driver}\"")
        else:
            out.append(l)
    if hits != 1:
        raise SystemExit(f"extract_ref_members: marker `{marker}` must occur exactly once in {driver} (found {hits})")
    sys.stdout.write(f"#line 1 \"{driver}\"\n" + "\n".join(out))
    print(f"extract_ref_members: {len(spans)} functions spliced into {driver}", file=sys.stderr)


if __name__ == "__main__":
    main()
