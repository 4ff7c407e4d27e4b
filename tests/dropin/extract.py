#!/usr/bin/env python
"""Pulls named test functions (and setUp / tearDown) out of one of the reference's own Unity test
files, WHERE IT LIES under /root/reference, into a scratch include file OUTSIDE the repository.

    python tests/dropin/extract.py <reference test .cpp> <out .inc> <function> [<function> ...]

The extracted text is the reference's, unmodified: tests/dropin/dropin_main.cpp compiles it against
the reference's own headers with -DSP_B200_USE_REFERENCE_TYPES and links libspb200.so instead of the
reference's bvh.cpp / sp_scene.cpp / sp_material_system.cpp / simd_path_tracer.cpp -- the drop-in
INTEGRATION.md describes, exercised by the reference's own assertions.  Nothing extracted is ever
written into the repository (the Makefile points <out> at $TMPDIR)."""
import re
import sys


def extract(text, name):
    m = re.search(r"^void\s+%s\s*\([^)]*\)\s*\{" % re.escape(name), text, re.M)
    if not m:
        raise SystemExit(f"{name}: not found")
    depth, i = 0, m.end() - 1
    while True:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[m.start():i + 1]
        i += 1


def main():
    src, out, names = sys.argv[1], sys.argv[2], sys.argv[3:]
    text = open(src).read()
    parts = [f"// extracted from {src} by tests/dropin/extract.py; not part of the repository\n"]
    for n in names:
        parts.append(extract(text, n))
    open(out, "w").write("\n\n".join(parts) + "\n")


if __name__ == "__main__":
    main()
