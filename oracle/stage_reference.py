"""Stage the UNMODIFIED reference under baseline/_ref/reference/ (git-ignored; travels to the GPU box with the snapshot).

   python oracle/stage_reference.py            # build container only: needs /root/reference (or $SINC_REFERENCE)

The reference is pure Python with no setup.py / pyproject: `pip install /root/reference` has nothing to build, so the
"install" is a verbatim copy of its source tree, byte for byte, into the one directory the task reserves for it
(baseline/_ref is listed in .gitignore -- nothing of the reference ever enters the history).  Test infrastructure:
  * bench.py --impl reference and the cpu_baseline leg time the REAL agents.perception.SincConv1d from this copy
    (kind "reference"); without it they fall back to the oracle port (kind "port");
  * tests/test_gpu_reference_agent.py and examples/cfg3_harness.py construct the real PerceptionAgent / CPEA / MSA /
    SpeechEnhancer from it on the GPU box, where /root/reference does not exist.
The product path never looks here.  A manifest with the sha256 of every staged file is written next to the copy so that
a test can tell the copy is still byte-identical to what was staged.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEST = os.path.join(ROOT, "baseline", "_ref", "reference")
KEEP_EXT = (".py", ".txt", ".md")


def staged_path():
    """baseline/_ref/reference if a staged copy with the hot path is present, else None."""
    return DEST if os.path.exists(os.path.join(DEST, "agents", "perception.py")) else None


def reference_root():
    """Where the real reference can be imported from right now: the source tree (build container) or the staged copy."""
    src = os.environ.get("SINC_REFERENCE", "/root/reference")
    if os.path.exists(os.path.join(src, "agents", "perception.py")):
        return src
    return staged_path()


def stage(src=None, verbose=True):
    src = src or os.environ.get("SINC_REFERENCE", "/root/reference")
    if not os.path.exists(os.path.join(src, "agents", "perception.py")):
        raise SystemExit(f"{src} does not hold the reference (agents/perception.py missing)")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d not in (".git", "__pycache__", ".pytest_cache")]
        for f in filenames:
            if not f.endswith(KEEP_EXT):
                continue
            rel = os.path.relpath(os.path.join(dirpath, f), src)
            out = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(dirpath, f), out)
            with open(out, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(os.path.dirname(DEST), "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files under {DEST}")
    return DEST


if __name__ == "__main__":
    stage(sys.argv[1] if len(sys.argv) > 1 else None)
