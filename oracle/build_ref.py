#!/usr/bin/env python
"""Build the reference's own Cython extensions into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The three native modules of SDFIdk/malstroem (malstroem/algorithms/speedups/_fill.pyx, _flow.pyx,
_label.pyx) are compiled from where they lie under /root/reference: a scratch copy is made under /tmp
(the reference tree is read-only and the build writes next to the sources), the dtype spellings that
Cython 3 / numpy 2 removed are patched there (np.int_t -> np.npy_long, np.int -> np.int64; SURVEY.md
F9), the extensions are built with the reference's own setup.py, and ONLY the resulting binaries are
copied to oracle/_ref/.  No reference source enters this repository; oracle/_ref/ is git-ignored and
travels to the GPU box with gpurun like any other built .so.

This runs only where /root/reference exists (the build container).  On the GPU box the prebuilt
binaries are used as they are.
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MALSTROEM_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(force=False):
    if not os.path.isdir(os.path.join(REF, "malstroem")):
        return False
    have = glob.glob(os.path.join(OUT, "_fill*.so")) and glob.glob(os.path.join(OUT, "_flow*.so")) \
        and glob.glob(os.path.join(OUT, "_label*.so"))
    if have and not force:
        return True
    tmp = tempfile.mkdtemp(prefix="malstroem_ref_build_")
    try:
        for name in ("malstroem", "setup.py", "README.rst", "README.md", "requirements.txt", "setup.cfg",
                     "MANIFEST.in"):
            src = os.path.join(REF, name)
            if os.path.isdir(src):
                shutil.copytree(src, os.path.join(tmp, name))
            elif os.path.exists(src):
                shutil.copy(src, tmp)
        sp = os.path.join(tmp, "malstroem", "algorithms", "speedups")
        for fn in ("_flow.pyx", "_label.pyx"):
            p = os.path.join(sp, fn)
            s = open(p).read()
            s = s.replace("np.int_t", "np.npy_long")
            s = re.sub(r"dtype=np\.int \)", "dtype=np.int64 )", s)
            s = re.sub(r"np\.int\)", "np.int64)", s)
            open(p, "w").write(s)
        env = dict(os.environ)
        env.setdefault("CFLAGS", "-O2")
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-4000:])
            raise RuntimeError("reference Cython build failed")
        os.makedirs(OUT, exist_ok=True)
        n = 0
        for so in glob.glob(os.path.join(sp, "*.so")):
            shutil.copy(so, OUT)
            n += 1
        if n != 3:
            raise RuntimeError("expected 3 reference extension binaries, got %d" % n)
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference tree not present; nothing built")
