#!/usr/bin/env python
"""Install the UNMODIFIED reference's two pure-Python packages (curdleproofs, merlin_transcripts) into baseline/_ref
(git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built .so files).

    python tools/install_reference.py [--force]

What happens, in order (recorded in DESIGN.md section 2):
  1. the contract's command  `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse
     --target baseline/_ref /root/reference/<pkg>`.  It fails in this image: both packages name poetry-core as their
     build backend, and poetry is neither importable nor in /opt/wheelhouse.
  2. fallback: the package directory is copied to a scratch directory under /tmp, ONLY pyproject.toml's [build-system]
     table is pointed at setuptools (packaging metadata; no source file is touched - the installed .py files are
     compared byte for byte with /root/reference afterwards), and the same pip command installs the copy.
The reference's arithmetic dependency (py_arkworks_bls12381, a Rust wheel) cannot be installed offline: at run time that
name is provided either by dropin/ (the B200 library) or by oracle/standin (the CPU oracle), chosen through PYTHONPATH.
"""
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DEST = os.path.join(ROOT, "baseline", "_ref")
PKGS = ("merlin_transcripts", "curdleproofs")

SETUPTOOLS_PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "%(name)s"
version = "%(version)s"

[tool.setuptools]
packages = ["%(name)s"]

[tool.setuptools.package-data]
"%(name)s" = ["py.typed"]
"""


def _pip(src):
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade", "--find-links", "/opt/wheelhouse",
           "--target", DEST, src]
    return subprocess.run(cmd, capture_output=True, text=True)


def _version(pyproject):
    for line in open(pyproject):
        if line.strip().startswith("version"):
            return line.split("=", 1)[1].strip().strip('"')
    return "0"


def installed():
    return all(os.path.isfile(os.path.join(DEST, p, "__init__.py")) for p in PKGS)


def install(force=False, quiet=False):
    """Returns a dict {package: how it was installed}.  No-op when already installed (or when /root/reference is
    absent: the GPU box uses what travelled with the snapshot)."""
    log = {}
    if installed() and not force:
        return {p: "present" for p in PKGS}
    if not os.path.isdir(REF):
        return {p: "reference not mounted" for p in PKGS}
    os.makedirs(DEST, exist_ok=True)
    for name in PKGS:
        src = os.path.join(REF, name)
        r = _pip(src)
        how = "pip (contract command)"
        if r.returncode != 0:
            tmp = tempfile.mkdtemp(prefix="cpg_ref_")
            try:
                work = os.path.join(tmp, name)
                shutil.copytree(src, work, ignore=shutil.ignore_patterns("dist", "*.lock", "__pycache__", ".mypy_cache"))
                version = _version(os.path.join(work, "pyproject.toml"))
                with open(os.path.join(work, "pyproject.toml"), "w") as f:
                    f.write(SETUPTOOLS_PYPROJECT % {"name": name, "version": version})
                r = _pip(work)
                how = "pip from a /tmp copy with [build-system] pointed at setuptools (poetry-core is not installable offline)"
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        if r.returncode != 0:
            raise RuntimeError("installing %s failed:\n%s" % (name, (r.stdout + r.stderr)[-2000:]))
        # the installed sources are the reference's, byte for byte
        pkg = os.path.join(src, name)
        for fn in sorted(os.listdir(pkg)):
            if fn.endswith(".py"):
                assert filecmp.cmp(os.path.join(pkg, fn), os.path.join(DEST, name, fn), shallow=False), "installed %s/%s differs from the reference" % (name, fn)
        log[name] = how
        if not quiet:
            print("%s -> %s: %s" % (name, os.path.relpath(DEST, ROOT), how))
    return log


if __name__ == "__main__":
    install(force="--force" in sys.argv)
