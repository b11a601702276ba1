#!/usr/bin/env python
"""Convert a SAMRAI-format CALPHAD data base of the reference
(thermodynamic_data/calphadAuNi.dat) into the JSON parameter record shipped in
ampe_b200/data/.  Run in the build container only (reads /root/reference):

    python tools/make_calphad_json.py /root/reference/thermodynamic_data/calphadAuNi.dat \
        ampe_b200/data/calphadAuNi.json

Only numbers are transcribed (parameters of the Newton kernel); no code.
"""
import json
import re
import sys


def parse(text):
    text = re.sub(r"//[^\n]*", "", text)
    root = {}
    stack = [root]
    for line in text.splitlines():
        line = line.strip()
        while line:
            if line.startswith("}"):
                stack.pop()
                line = line[1:].strip()
                continue
            m = re.match(r"^(\w+)\s*\{(.*)$", line)
            if m:
                d = {}
                stack[-1][m.group(1)] = d
                stack.append(d)
                line = m.group(2).strip()
                continue
            m = re.match(r"^(\w+)\s*=\s*([^}]*)(.*)$", line)
            if m:
                v = m.group(2).strip()
                if v.startswith('"'):
                    stack[-1][m.group(1)] = v.strip('"')
                else:
                    stack[-1][m.group(1)] = [float(x) for x in v.replace(",", " ").split()]
                line = m.group(3).strip()
                continue
            raise ValueError("cannot parse: " + line)
    return root


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    db = parse(open(src).read())
    json.dump(db, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst)
