#!/usr/bin/env python
"""Stage the UNMODIFIED reference tree where a GPU box can see it (TEST INFRASTRUCTURE ONLY).

    python oracle/stage_reference.py [--remove]

`/root/reference` exists in the build container only.  SURVEY.md App. C's recipe — a verbatim copy under the git-ignored
`baseline/_ref/` — lets the copy travel with the repository snapshot, so that on a GPU box
  * `tests/test_gpu_reference_learn.py` can drive the reference's own `learn()` on the fused kernels, and
  * `bench.py --impl reference` can time the Python original (DummyVecEnv, SubprocVecEnv, collect_rollouts, TD3.train) beside its C port.
Nothing in the product reads it (`oracle/refload.py` is the only loader, used by tests, the oracle scripts and the reference arm), it is
never committed (`.gitignore`), and no file of it is edited: the missing `core/version.txt` is supplied by `refload`'s scratch mirror.
"""
import argparse
import os
import shutil
import stat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CSTR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def _writable(path):
    for d, _, files in os.walk(path):
        os.chmod(d, os.stat(d).st_mode | stat.S_IWUSR)
        for f in files:
            p = os.path.join(d, f)
            os.chmod(p, os.stat(p).st_mode | stat.S_IWUSR)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--remove", action="store_true", help="delete the staged copy")
    args = ap.parse_args()
    if os.path.isdir(DST):
        _writable(DST)
        shutil.rmtree(DST)
    if args.remove:
        print(f"removed {DST}")
        return
    if not os.path.isfile(os.path.join(SRC, "twoseriescstr.py")):
        raise SystemExit(f"no reference tree at {SRC}")
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".git"))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"staged {n} files of {SRC} under {DST} (git-ignored)")


if __name__ == "__main__":
    main()
