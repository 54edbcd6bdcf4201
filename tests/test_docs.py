"""CPU: the documents only cite evidence that is in the repository (every `profiles/...`, `tools/...`, `tests/...`,
`shim/...`, `oracle/...`, `gpu_ai_b200/...`, `include/...` path and every `rNN*` artefact name in back-ticks resolves)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOCS = ("DESIGN.md", "BASELINE.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md"))
BUILT = ("oracle/_ref", "shim/_ref", "gpu_ai_b200/libb2p", "oracle/liboracle", "baseline/_ref")   # build products, git-ignored


def _exists(path):
    path = path.rstrip(".,")
    return bool(glob.glob(os.path.join(ROOT, path)) or glob.glob(os.path.join(ROOT, path + "*")))


def test_cited_files_exist():
    missing = []
    for doc in DOCS:
        text = open(os.path.join(ROOT, doc)).read()
        for m in re.findall(r"`((?:profiles|tools|tests|shim|oracle|gpu_ai_b200|include)/[A-Za-z0-9_./*{},\-]+)`", text):
            names = [m]
            if "{" in m and "}" in m:
                pre, rest = m.split("{", 1)
                mid, post = rest.split("}", 1)
                names = [pre + x + post for x in mid.split(",")]
            for n in names:
                n = n.split("::")[0]
                if n.startswith(BUILT):
                    continue
                if not _exists(n):
                    missing.append((doc, n))
        for m in re.findall(r"`(r0[0-9][a-z_0-9]*[A-Za-z0-9_.*\-]*)`", text):
            if not _exists(os.path.join("profiles", m)):
                missing.append((doc, "profiles/" + m))
    assert not missing, missing
