"""Places the UNMODIFIED hot-path files of the reference under baseline/_ref/ (git-ignored; travels to the GPU box with the
snapshot) so that `bench.py --impl reference` and the reference-driven tests can run the reference's own code there through
the fairseq stand-ins of oracle/ref_shim.  The reference is pure Python without a setup.py, so "installing" it is copying the
files it needs for this path, byte for byte; nothing is copied into tracked source.

    python tools/install_reference.py [--src /root/reference]
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/__init__.py", "models/search.py", "models/sequence_generator.py", "models/ofa/__init__.py", "models/ofa/ofa.py",
         "models/ofa/unify_transformer.py", "models/ofa/unify_transformer_layer.py", "models/ofa/unify_multihead_attention.py",
         "models/ofa/resnet.py", "models/ofa/frozen_bn.py", "criterions/label_smoothed_cross_entropy.py", "data/data_utils.py",
         "data/__init__.py", "utils/trie.py", "LICENSE"]


def install(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "models", "ofa")):
        return False
    for f in FILES:
        s, d = os.path.join(src, f), os.path.join(DST, f)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--src") + 1] if "--src" in sys.argv else "/root/reference"
    print("installed" if install(src) else "reference tree not found at %s" % src, "->", DST)
