#!/usr/bin/env python3
"""Generate tests/golden/ fixtures from the mounted reference tree.

Run in the build container only (needs /root/reference); the outputs are
committed so the GPU box never reads the reference:

  captures.npz      the three 512 KiB IQ captures of test_iq/ (the reference's own
                    test inputs, tests/test.rs:21,34,48), losslessly re-encoded:
                    int16 values in FILE order [im0, re0, im1, re1, ...]
                    (utils.rs:29-31 reads im first), zlib-compressed by numpy.
  golden_frames.json the expected frame bytes of tests/test.rs:22-28, :35-42,
                    :49-56, parsed from the hex!() literals, with their line numbers.
"""
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> int:
    if not os.path.isdir(REF):
        print("reference tree not mounted; fixtures are already committed", file=sys.stderr)
        return 1
    caps = {}
    for fn in sorted(os.listdir(os.path.join(REF, "test_iq"))):
        if fn.endswith(".iq"):
            raw = np.fromfile(os.path.join(REF, "test_iq", fn), dtype="<i2")
            assert raw.size == 2 * 0x20000, (fn, raw.size)
            caps[fn[:-3]] = raw
    np.savez_compressed(os.path.join(HERE, "captures.npz"), **caps)

    src = open(os.path.join(REF, "tests", "test.rs")).read().splitlines()
    tests, cur = {}, None
    for ln, line in enumerate(src, 1):
        m = re.search(r'let filename = "test_iq/(test_\d+)\.iq"', line)
        if m:
            cur = m.group(1)
            tests[cur] = []
        m = re.search(r'hex!\("([0-9a-fA-F]+)"\)', line)
        if m and cur:
            tests[cur].append({"hex": m.group(1).lower(), "line": ln})
    with open(os.path.join(HERE, "golden_frames.json"), "w") as f:
        json.dump({"source": "tests/test.rs", "captures": tests}, f, indent=1)
    print({k: len(v) for k, v in tests.items()}, {k: v.size for k, v in caps.items()})
    return 0


if __name__ == "__main__":
    sys.exit(main())
