#!/usr/bin/env python
"""Proves that a refactor left already-validated kernels untouched: compares two builds of libspruce_b200.so function by function
at the SASS level (instruction text + encoding, addresses ignored) and prints the static size / FP64 count of every stage-kernel
instance of the new build.

    cuobjdump -sass old/libspruce_b200.so > /tmp/old.sass      # or pass the .so files themselves
    python scripts/sass_identity.py /tmp/old.sass spruce_b200/lib/libspruce_b200.so
    python scripts/sass_identity.py profiles/r1_validated_kernels.sass.gz spruce_b200/lib/libspruce_b200.so     # against the build the round-1 GPU runs validated

A trailing default template argument that the new build added (`..., 0>` mangled as `ELi0EEEv`) is ignored when names are matched.
Exit status 1 when any function present in both builds differs.  No GPU needed."""
import re
import subprocess
import sys


def sass_lines(path):
    if path.endswith(".so"):
        return subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout.splitlines()
    if path.endswith(".gz"):
        import gzip
        return gzip.open(path, "rt").read().splitlines()
    return open(path).read().splitlines()


def functions(path):
    out, name = {}, None
    for line in sass_lines(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1); out[name] = []
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            out[name].append(re.sub(r"^\s+/\*[0-9a-f]+\*/", "", line).strip())
    return out


def key(n):
    return re.sub(r"ELi0EEEv", "EEEv", n)


def main():
    old, new = functions(sys.argv[1]), functions(sys.argv[2])
    newk = {key(n): v for n, v in new.items()}
    same = diff = 0
    for n, v in old.items():
        w = newk.get(key(n))
        if w is None:
            print("only in old:", n); continue
        if v == w:
            same += 1
        else:
            diff += 1
            k = next((i for i, (x, y) in enumerate(zip(v, w)) if x != y), min(len(v), len(w)))
            print("DIFFERS: %s (%d vs %d instructions, first difference at #%d)" % (n, len(v), len(w), k))
    print("identical: %d   different: %d   new functions: %d" % (same, diff, len(new) - same - diff))
    for n, v in sorted(new.items()):
        if "k_mhd_stage_xy" in n:
            fp = sum(1 for i in v if re.match(r"(@!?U?P\d+ )?D(FMA|MUL|ADD|SETP)", i))
            print("  %-110s %5d instructions, %4d FP64" % (n, len(v), fp))
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())
