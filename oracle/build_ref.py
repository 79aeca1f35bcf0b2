"""Install the UNMODIFIED reference into ``oracle/_ref`` (git-ignored; it travels to the GPU box with the push, like the
built ``.so`` files).  TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.

    python oracle/build_ref.py            # needs /root/reference (build container only)

The reference is pure Python, so "building" it is an install: the package tree ``src/earthkit/meteo`` is placed under
``oracle/_ref/site/`` byte for byte (the same result as ``pip install --no-deps --target``, without running the reference's
own build system, which needs setuptools_scm and git metadata), together with the reference's thermo tests and the golden
CSVs they read (``tests/thermo/test_thermo.py``, ``tests/data/*.csv``) under ``oracle/_ref/tests/``.  Nothing is copied
into the tracked tree.  ``oracle/_ref/MANIFEST.json`` records the sha256 of every installed file next to the sha256 of
its source, so "unmodified" can be checked later without the source tree.

The only import the reference's thermo path needs that is not installable offline is the third-party
``earthkit-utils`` (pinned ``>=0.2`` in the reference's pyproject.toml); the numpy stand-in ``oracle/refshim`` provides
it (see its docstring).  ``bench.py --impl reference`` puts ``oracle/_ref/site`` and ``oracle/refshim`` on ``sys.path``,
runs the reference's own 92 thermo tests against the installed copy, and then times
``earthkit.meteo.thermo.array`` through the reference's public functions.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EK_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _install_tree(src, dst, manifest, keep=lambda name: True):
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, src)
        for name in sorted(filenames):
            if name.endswith((".pyc", ".pyo")) or not keep(name):
                continue
            s = os.path.join(dirpath, name)
            d = os.path.normpath(os.path.join(dst, rel, name))
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[os.path.relpath(d, DEST)] = {"source": os.path.relpath(s, REF), "sha256": _sha(d), "source_sha256": _sha(s)}


def available():
    return os.path.isdir(os.path.join(REF, "src", "earthkit", "meteo"))


def installed():
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def build(verbose=True):
    if not available():
        if verbose:
            print(f"build_ref: {REF} not present; keeping the existing oracle/_ref ({'present' if installed() else 'absent'})")
        return installed()
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    _install_tree(os.path.join(REF, "src", "earthkit", "meteo"), os.path.join(DEST, "site", "earthkit", "meteo"), manifest)
    _install_tree(os.path.join(REF, "tests", "thermo"), os.path.join(DEST, "tests", "thermo"), manifest)
    _install_tree(os.path.join(REF, "tests", "data"), os.path.join(DEST, "tests", "data"), manifest, keep=lambda n: n.endswith(".csv"))
    assert all(v["sha256"] == v["source_sha256"] for v in manifest.values())
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"reference_root": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"build_ref: installed {len(manifest)} files of the unmodified reference into {DEST}")
    return True


def sys_path_entries():
    """What has to be on sys.path to import the installed reference: its site dir and the earthkit-utils stand-in."""
    return [os.path.join(DEST, "site"), os.path.join(HERE, "refshim")]


def verify():
    """Every installed file still has the sha256 recorded at install time (i.e. is the reference's own byte stream)."""
    with open(os.path.join(DEST, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    bad = [rel for rel, v in files.items() if _sha(os.path.join(DEST, rel)) != v["source_sha256"]]
    return len(files), bad


def run_reference_tests(timeout=600):
    """The reference's own thermo tests against the installed copy, in a fresh interpreter.  Returns (ok, summary line)."""
    import subprocess

    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(sys_path_entries() + [env.get("PYTHONPATH", "")])
    cmd = [sys.executable, "-m", "pytest", os.path.join(DEST, "tests", "thermo", "test_thermo.py"), "-q", "-p", "no:cacheprovider",
           "-o", "addopts=", "--rootdir", os.path.join(DEST, "tests"), "-W", "ignore"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=os.path.join(DEST, "tests"))
    tail = [ln for ln in r.stdout.strip().splitlines() if ln.strip()]
    return r.returncode == 0, (tail[-1] if tail else r.stderr.strip()[-200:])


if __name__ == "__main__":
    ok = build()
    if ok and "--test" in sys.argv:
        print(run_reference_tests())
    sys.exit(0 if ok else 1)
