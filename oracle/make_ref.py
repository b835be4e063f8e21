"""Recipe for oracle/_ref/: the UNMODIFIED reference package, copied file by file from /root/reference so
that ``bench.py --impl reference`` can import and time the reference's own replay + learner code on the GPU
box (where /root/reference does not exist).  Test/bench infrastructure only.

  python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

Outputs go to oracle/_ref/ only, which is git-ignored (reference sources never enter the history) but not
gpurun-ignored, so the copy travels to the GPU box with the snapshot like the built .so files.  Nothing is
patched: the files are byte-identical copies (a manifest with their sha256 is written next to them); the
three third-party modules the reference imports and this image lacks (lz4, prefetch_generator, gymnasium via
agent0.common.atari_wrappers) are supplied as sys.modules stubs by oracle/ref_arm.py at import time.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
# the path of SURVEY section 8: replay, the learners, the trainer loop, the DataLoader pump; plus what they import
FILES = ["agent0/__init__.py", "agent0/deepq/__init__.py", "agent0/deepq/replay.py", "agent0/deepq/agent.py",
         "agent0/deepq/trainer.py", "agent0/deepq/config.py", "agent0/deepq/model.py", "agent0/common/__init__.py",
         "agent0/common/utils.py"]


def make(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "agent0")):
        if verbose:
            print(f"{SRC} is not present: oracle/_ref/ left as it is")
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()              # namespace package in the reference
            manifest[rel] = "(empty: no such file in the reference)"
        else:
            raise FileNotFoundError(src)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files copied from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
