"""Installs the UNMODIFIED reference (DegardinBruno/Kinetic-GAN, a script tree without packaging metadata) into baseline/_ref/
so that it travels to the GPU box with the snapshot (baseline/_ref is git-ignored, not gpurun-ignored):

    python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference + setup.py>

The reference has no setup.py / pyproject.toml, and /root/reference is read-only, so the install runs from a copy under /tmp
to which ONLY a generated setup.py is added (package list: `models`, `models.init_gan`; not a single reference source line is
touched).  Called by __graft_entry__.build() when /root/reference exists; a no-op elsewhere (the GPU box uses what was built).
bench.py --impl reference / the `gpu_reference` leg import the installed modules through oracle/ref_runner.py."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KGAN_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")

SETUP = '''from setuptools import setup
setup(name="kinetic-gan-reference", version="0", description="unmodified DegardinBruno/Kinetic-GAN model sources",
      packages=["models", "models.init_gan"])
'''


def install(force=False):
    """-> (ok, one-line outcome)."""
    marker = os.path.join(TARGET, "models", "generator.py")
    if os.path.exists(marker) and not force:
        return True, "already installed"
    if not os.path.isdir(os.path.join(REF, "models")):
        return False, "reference tree not present at %s" % REF
    tmp = tempfile.mkdtemp(prefix="kgan_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns("__pycache__", "*.gif", ".git"))
        with open(os.path.join(src, "setup.py"), "w") as f:
            f.write(SETUP)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", TARGET, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(marker):
            return False, "pip install failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:200]
        return True, "installed with pip --target baseline/_ref (setup.py generated; sources unmodified)"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    ok, msg = install(force="--force" in sys.argv)
    print(("ok: " if ok else "FAILED: ") + msg)
    sys.exit(0 if ok else 1)
