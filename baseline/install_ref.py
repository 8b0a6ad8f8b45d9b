"""Copies the UNMODIFIED reference env into git-ignored baseline/_ref/ so that it travels to the GPU box.

    python baseline/install_ref.py [--reference /root/reference]

Route: `python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target
baseline/_ref <copy of the reference under /tmp>` (the source tree is read-only and setup.py builds in place, hence the
copy; --no-deps because setup.py pins gym==0.21.0 / scipy==1.7.3 / numpy==1.21.* / stable_baselines3, none of which
is in the offline wheelhouse — dependency resolution is the only thing that fails).  pip installs the `my_environment`
package (envs, utils, wrappers); `config.yaml` (not package data) is copied beside it.  If pip is unusable the same
files are copied directly.  baseline/_ref/ is listed in .gitignore (reference sources never enter the history) but
NOT in .gpurunignore, so it ships to the GPU box.
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "config.yaml",
    "my_environment/__init__.py",
    "my_environment/envs/__init__.py",
    "my_environment/envs/rocket_env.py",
    "my_environment/utils/simulator.py",
    "my_environment/wrappers/__init__.py",
    "my_environment/wrappers/wrappers.py",
]


def install(reference="/root/reference", dest=DEST):
    """Returns the manifest {relative path: sha256}; raises if the reference tree is not there."""
    if not os.path.isfile(os.path.join(reference, "my_environment", "envs", "rocket_env.py")):
        raise FileNotFoundError(f"reference tree not found at {reference}")
    how = "pip --no-deps --target"
    shutil.rmtree(dest, ignore_errors=True)
    with tempfile.TemporaryDirectory() as tmp:
        src_copy = os.path.join(tmp, "reference")
        shutil.copytree(reference, src_copy, ignore=shutil.ignore_patterns("*.zip", "*.png", ".git", "__pycache__"))
        res = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                              "--find-links", "/opt/wheelhouse", "--target", dest, src_copy],
                             capture_output=True, text=True)
    if res.returncode != 0 or not all(os.path.isfile(os.path.join(dest, r)) for r in FILES if r != "config.yaml"):
        how = "file copy (pip failed: %s)" % (res.stderr.strip().splitlines()[-1:] or ["?"])[0][:120]
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst):
            shutil.copyfile(src, dst)
        with open(src, "rb") as f0, open(dst, "rb") as f1:
            if f0.read() != f1.read():
                raise RuntimeError(f"{rel}: installed copy differs from the reference source")
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference, "how": how, "files": manifest}, f, indent=1)
    return manifest


def installed(dest=DEST):
    return os.path.isfile(os.path.join(dest, "my_environment", "envs", "rocket_env.py"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    m = install(a.reference)
    print(f"installed {len(m)} reference files into {DEST}")
