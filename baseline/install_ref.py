#!/usr/bin/env python
"""Install the UNMODIFIED reference package into baseline/_ref/ (TEST / BASELINE INFRASTRUCTURE ONLY).

`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <staged copy>`: the reference tree is
read-only and its build writes next to the sources, so a scratch copy under /tmp is installed; the only edit to
that copy is the 7-line dtype-spelling patch Cython 3 / numpy 2 require (np.int_t -> np.npy_long, np.int ->
np.int64; SURVEY.md F9, same as oracle/build_ref.py).  baseline/_ref/ is git-ignored and travels to the GPU box
with gpurun.  The tool-layer tests (tests/test_tool_layer.py) put it on sys.path and drive the reference's own
DemTool / BluespotTool / StreamTool / RainTool on top of malstroem_b200.speedups.enable().

Runs only where /root/reference exists (the build container)."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MALSTROEM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def install(force=False):
    if not os.path.isdir(os.path.join(REF, "malstroem")):
        return False
    if os.path.isdir(os.path.join(OUT, "malstroem")) and not force:
        return True
    tmp = tempfile.mkdtemp(prefix="malstroem_ref_install_")
    try:
        for name in os.listdir(REF):
            if name.startswith(".") or name in ("tests", "docs"):
                continue
            src = os.path.join(REF, name)
            if os.path.isdir(src):
                shutil.copytree(src, os.path.join(tmp, name))
            else:
                shutil.copy(src, tmp)
        sp = os.path.join(tmp, "malstroem", "algorithms", "speedups")
        for fn in ("_flow.pyx", "_label.pyx"):
            p = os.path.join(sp, fn)
            s = open(p).read()
            s = s.replace("np.int_t", "np.npy_long")
            s = re.sub(r"dtype=np\.int \)", "dtype=np.int64 )", s)
            s = re.sub(r"np\.int\)", "np.int64)", s)
            open(p, "w").write(s)
        shutil.rmtree(OUT, ignore_errors=True)
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", OUT, tmp],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-4000:])
            raise RuntimeError("pip install of the reference failed")
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "installed" if ok else "reference tree not present; nothing installed")
