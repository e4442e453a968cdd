"""TEST / BASELINE INFRASTRUCTURE — recipe that stages the UNMODIFIED reference for the GPU box.

    python -m oracle.make_ref

Copies the hot-path packages of /root/reference (agents/, envs/, models/, utils.py and the yaml files the named
configurations use) byte for byte into the git-ignored ``oracle/_ref/pyref/`` so that ``bench.py --impl reference`` can time
the reference's own torch CPU path on the GPU box's host cores (the reference tree itself does not exist there; SURVEY.md
§8(d) "CPU baseline timing").  Nothing is copied into the repository's history: ``oracle/_ref/`` is listed in .gitignore
(but not in .gpurunignore, so it travels with the snapshot like the built .so files).  The three third-party modules the
reference imports and this image lacks (gym 0.17.3, seaborn, ConfigSpace) come from ``oracle/stubs`` as in the tests.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "pyref")
PACKAGES = ["agents", "envs", "models"]
FILES = ["utils.py", "__init__.py", "default_config_cartpole_syn_env.yaml", "default_config_acrobot_syn_env.yaml",
         "default_config_cartpole_reward_env.yaml", "default_config_acrobot.yaml", "default_config_cartpole.yaml"]


def available():
    return os.path.isdir(os.path.join(SRC, "agents"))


def stage(force=False):
    """Returns the staged path (or None when /root/reference is absent and nothing was staged before)."""
    stamp = os.path.join(DST, "MANIFEST.json")
    if not available():
        return DST if os.path.isfile(stamp) else None
    if os.path.isfile(stamp) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for p in PACKAGES:
        shutil.copytree(os.path.join(SRC, p), os.path.join(DST, p), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in FILES:
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            path = os.path.join(root, f)
            with open(path, "rb") as fh:
                manifest[os.path.relpath(path, DST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(stamp, "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(stage(force=True))
