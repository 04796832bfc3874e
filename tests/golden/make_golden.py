#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference's own test data (run in the build container only).

The reference (ianic/flate) keeps its golden vectors as binary fixtures under
src/flate/testdata/ plus token lists written as Zig source in testdata/block_writer.zig.
This script copies the binary fixtures verbatim (they are test vectors, not source code) and
converts the Zig token lists into a neutral JSON form:

    block_writer_tokens.json : [{"input": name|"" , "want": pattern|"" , "want_no_input": pattern,
                                 "tokens": [[lit] | [dist, len], ...]}, ...]

Numeric goldens (token counts, compressed sizes, fuzz error classes) are small enough that the
tests state them inline with a file:line citation.
"""
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
TD = os.path.join(REF, "src", "flate", "testdata")


def parse_tokens(body):
    toks = []
    for m in re.finditer(r"L\('(\\.|[^\\])'\)|L\((0x[0-9a-fA-F]+|\d+)\)|M\((\d+),\s*(\d+)\)|\bml\b", body):
        if m.group(1) is not None:
            ch = m.group(1)
            if ch.startswith("\\"):
                ch = {"\\n": "\n", "\\t": "\t", "\\r": "\r", "\\'": "'", "\\\\": "\\"}[ch]
            toks.append([ord(ch)])
        elif m.group(2) is not None:
            toks.append([int(m.group(2), 0)])
        elif m.group(3) is not None:
            toks.append([int(m.group(3)), int(m.group(4))])
        else:
            toks.append([1, 258])  # ml = M(1, 258)
    return toks


def main():
    for sub in ("block_writer", "fuzz"):
        dst = os.path.join(HERE, sub)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(TD, sub))):
            shutil.copyfile(os.path.join(TD, sub, f), os.path.join(dst, f))
    shutil.copyfile(os.path.join(TD, "rfc1951.txt"), os.path.join(HERE, "rfc1951.txt"))

    src = open(os.path.join(TD, "block_writer.zig")).read()
    src = src[src.index("break :blk"):]
    cases = []
    for m in re.finditer(r"TestCase\{(.*?)\n        \},", src, re.S):
        body = m.group(1)
        def field(name):
            fm = re.search(r"\.%s = \"([^\"]*)\"" % name, body)
            return fm.group(1) if fm else ""
        tb = body[body.index(".tokens"):]
        cases.append({"input": field("input"), "want": field("want"),
                      "want_no_input": field("want_no_input"), "tokens": parse_tokens(tb)})
    json.dump(cases, open(os.path.join(HERE, "block_writer_tokens.json"), "w"))
    print("cases:", len(cases), [len(c["tokens"]) for c in cases])


if __name__ == "__main__":
    main()
