"""Stage the UNMODIFIED upstream reference for the GPU box:  python tools/stage_ref.py

Copies ``/root/reference/src/rl8`` and the bundled example envs (``examples/*/env.py``) into the git-ignored
``oracle/_ref/`` (listed in .gitignore, NOT in .gpurunignore, so it travels with the ``gpurun`` snapshot like the
built ``.so``).  ``bench.py --impl reference`` then times the reference's own ``AlgorithmConfig(...).build(env)``
/ ``collect()`` / ``step()`` on the box's host cores (``cpu_baseline.kind = "reference"``) behind the
``oracle/refshim`` stand-ins for tensordict / torchrl / mlflow (which carry no arithmetic).
Nothing is copied into tracked paths; without ``/root/reference`` (the GPU box) this script is a no-op and the
bench falls back to the CPU port (``oracle/ppo_oracle.py``, ``kind = "port"``).
"""

from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")


def stage() -> bool:
    if not os.path.isdir(os.path.join(SRC, "src", "rl8")):
        return os.path.isdir(os.path.join(DST, "src", "rl8"))
    shutil.rmtree(DST, ignore_errors=True)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    shutil.copytree(os.path.join(SRC, "src", "rl8"), os.path.join(DST, "src", "rl8"), ignore=ignore)
    ex_src, ex_dst = os.path.join(SRC, "examples"), os.path.join(DST, "examples")
    os.makedirs(ex_dst, exist_ok=True)
    for name in sorted(os.listdir(ex_src)):
        env_py = os.path.join(ex_src, name, "env.py")
        if os.path.isfile(env_py):
            os.makedirs(os.path.join(ex_dst, name), exist_ok=True)
            shutil.copy2(env_py, os.path.join(ex_dst, name, "env.py"))
            init = os.path.join(ex_src, name, "__init__.py")
            if os.path.isfile(init):
                shutil.copy2(init, os.path.join(ex_dst, name, "__init__.py"))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC} (unmodified copy made by tools/stage_ref.py; git-ignored)\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else "reference not available", DST)
    sys.exit(0)
