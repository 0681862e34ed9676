"""Builds libhwer_b200 with an alternative score_filter.cu into variants/libhwer_b200_<name>.so (for same-box A/B
runs with scripts/ab_rounds.py).  Usage: python scripts/build_variant.py <name> <path/to/file.cu | git-rev> [file.cu]
(the csrc file that is swapped, default score_filter.cu)"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hybrid-weighted-embedding-recommender_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402

name, src = sys.argv[1], sys.argv[2]
which = sys.argv[3] if len(sys.argv) > 3 else "score_filter.cu"
out_dir = os.path.join(ROOT, "variants")
os.makedirs(out_dir, exist_ok=True)
tmp = tempfile.mkdtemp()
csrc = os.path.join(tmp, "pkg", "csrc")          # csrc includes "../../include/hwer_b200.h"
shutil.copytree(B.CSRC, csrc)
shutil.copytree(os.path.join(ROOT, "include"), os.path.join(tmp, "include"))
dst = os.path.join(csrc, which)
if os.path.exists(src):
    shutil.copy(src, dst)
else:
    rel = "hybrid-weighted-embedding-recommender_b200/csrc/" + which
    open(dst, "w").write(subprocess.run(["git", "show", "%s:%s" % (src, rel)], cwd=ROOT, capture_output=True, text=True,
                                        check=True).stdout)
out = os.path.join(out_dir, "libhwer_b200_%s.so" % name)
cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-o", out] + [os.path.join(csrc, s) for s in B.SOURCES]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
print(out)
