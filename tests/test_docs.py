"""The documents point at evidence and code by path: every `profiles/...`, `tools/...`, `tests/...`, `gspn_b200/...`, `oracle/...`,
`include/...` path written in backticks in DESIGN.md, README.md, INTEGRATION.md and profiles/README.md has to exist."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ["DESIGN.md", "README.md", "INTEGRATION.md", "profiles/README.md", "oracle/README.md"]
PREFIXES = ("profiles/", "tools/", "tests/", "gspn_b200/", "oracle/", "include/")


def _paths(text, base):
    for tok in re.findall(r"`([^`\s]+)`", text):
        tok = tok.rstrip(".,;:)")
        tok = re.sub(r":\d+(-\d+)?$", "", tok)  # file:line citations
        if tok.startswith(PREFIXES):
            yield tok
        elif base == "profiles" and re.match(r"^r0\d[\w{},.*|-]*\.(json|txt|csv)$", tok):
            yield "profiles/" + tok
        elif base == "profiles" and tok == "ncu_traffic.json":
            yield "profiles/" + tok


def _expand(tok):
    """`a_{x,y}_b.json` and `cfg3|cfg4` style shorthands -> concrete names; `*` -> glob."""
    m = re.search(r"\{([^{}]*)\}", tok)
    if m:
        out = []
        for alt in m.group(1).split(","):
            out += _expand(tok[:m.start()] + alt + tok[m.end():])
        return out
    return [tok]


def test_every_path_the_documents_cite_exists():
    missing = []
    for doc in DOCS:
        path = os.path.join(ROOT, doc)
        if not os.path.exists(path):
            continue
        base = os.path.dirname(doc)
        for tok in set(_paths(open(path).read(), base)):
            for name in _expand(tok):
                if "|" in name or "<" in name or "…" in name:
                    continue
                full = os.path.join(ROOT, name)
                if "*" in name:
                    ok = bool(glob.glob(full))
                elif name.startswith("oracle/_ref"):
                    ok = True  # built from /root/reference where that exists; git-ignored
                elif "::" in name:  # a test cited as file::function
                    f, fn = name.split("::", 1)
                    fp = os.path.join(ROOT, f)
                    ok = os.path.exists(fp) and ("def %s(" % fn.split("[")[0]) in open(fp).read()
                else:
                    ok = os.path.exists(full) or os.path.exists(full.rstrip("/"))
                if not ok:
                    missing.append("%s -> %s" % (doc, name))
    assert not missing, "\n".join(sorted(missing))
