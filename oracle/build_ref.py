"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Recipe for oracle/_ref: compiles the UNMODIFIED reference (every module under /root/reference/src, where the sources lie)
to CPython bytecode and stages the result -- bytecode only, no source text -- as ONE archive, oracle/_ref/reference_pyc.zip
(imported through zipimport; loose .pyc files do not survive the copy to the GPU box), which also carries the two data files
per model variant the reference's own from_pretrained calls read (vocab.txt, config.json of yaml/<variant>/).

    python -m oracle.build_ref          # in the build container; __graft_entry__.build() runs it when /root/reference exists

oracle/_ref is git-ignored (it stays out of history) and NOT gpurun-ignored: like the built .so it travels to the GPU box,
where /root/reference does not exist. oracle/ref_loader.py then imports the reference from it (sourceless import from the
archive), which is what lets `bench.py --impl reference` and the cpu_baseline leg time the reference's own code on the
box's host cores (cpu_baseline.kind "reference") instead of the restatement in oracle/port.py.

The bytecode is tied to this interpreter's magic number; the GPU box runs the same image. ref_loader checks the number
recorded in MANIFEST.json and ignores a stale tree.
"""
import hashlib
import importlib.util
import json
import os
import py_compile
import shutil
import sys
import warnings
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("VITCAP_REFERENCE_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ARCHIVE = "reference_pyc.zip"
DATA_FILES = ("vocab.txt", "config.json")


def magic_hex():
    return importlib.util.MAGIC_NUMBER.hex()


def build(verbose=False):
    """Returns the manifest dict, or None when the reference tree is absent (the GPU box: the staged tree is used as is)."""
    src = os.path.join(SRC_ROOT, "src")
    if not os.path.isdir(src):
        return None
    tmp = OUT + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    pyc = os.path.join(tmp, "_pyc")
    n_mod, digest = 0, hashlib.sha256()
    for root, dirs, files in os.walk(src):
        dirs.sort()
        rel = os.path.relpath(root, SRC_ROOT)
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s = os.path.join(root, f)
            d = os.path.join(pyc, rel, f + "c")          # legacy layout: module.pyc where module.py would be
            os.makedirs(os.path.dirname(d), exist_ok=True)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", SyntaxWarning)
                    py_compile.compile(s, cfile=d, dfile=os.path.join("<reference>", rel, f), doraise=True,
                                       invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            except py_compile.PyCompileError as e:       # python-2 era files the hot path never imports
                if verbose:
                    print("skipped (does not compile):", os.path.join(rel, f), "--", str(e).splitlines()[-1])
                continue
            with open(s, "rb") as fh:
                digest.update(hashlib.sha256(fh.read()).digest())
            n_mod += 1
    with zipfile.ZipFile(os.path.join(tmp, ARCHIVE), "w", zipfile.ZIP_DEFLATED) as z:
        for root, dirs, files in os.walk(pyc):
            dirs.sort()
            if root != pyc:                              # explicit directory entries: zipimport finds namespace packages
                z.write(root, os.path.relpath(root, pyc) + "/")     # (directories without __init__) only through them
            for f in sorted(files):
                full = os.path.join(root, f)
                z.write(full, os.path.relpath(full, pyc))
        # the two data files per model variant the reference's from_pretrained calls read ride in the same archive
        # (ref_loader extracts them to a temporary directory when it builds a model)
        ydir = os.path.join(SRC_ROOT, "yaml")
        variants = []
        for v in sorted(os.listdir(ydir)) if os.path.isdir(ydir) else []:
            if all(os.path.isfile(os.path.join(ydir, v, f)) for f in DATA_FILES):
                for f in DATA_FILES:
                    z.write(os.path.join(ydir, v, f), "yaml/%s/%s" % (v, f))
                variants.append(v)
    shutil.rmtree(pyc)
    man = {"what": "CPython bytecode of the unmodified reference's src/ tree (no source text; %s) + vocab.txt / config.json "
                   "of its model variants; built by oracle/build_ref.py" % ARCHIVE,
           "python_magic": magic_hex(), "python": sys.version.split()[0], "modules": n_mod, "variants": variants,
           "sources_sha256": digest.hexdigest()}
    with open(os.path.join(tmp, "MANIFEST.json"), "w") as fh:
        json.dump(man, fh, indent=1)
    shutil.rmtree(OUT, ignore_errors=True)
    os.replace(tmp, OUT)
    return man


def staged_root():
    """oracle/_ref when it holds a tree this interpreter can import, else None."""
    try:
        with open(os.path.join(OUT, "MANIFEST.json")) as fh:
            man = json.load(fh)
    except (OSError, ValueError):
        return None
    if man.get("python_magic") != magic_hex() or not os.path.isfile(os.path.join(OUT, ARCHIVE)):
        return None
    return OUT


def data_dir(root, variant_dir):
    """Directory holding vocab.txt / config.json of a model variant: yaml/<variant> of a source tree, or a temporary directory
    the two files are extracted to from the staged archive."""
    arc = os.path.join(root, ARCHIVE)
    if not os.path.isfile(arc):
        return os.path.join(root, "yaml", variant_dir)
    import tempfile
    out = tempfile.mkdtemp(prefix="vitcap_ref_%s_" % variant_dir[-6:])
    with zipfile.ZipFile(arc) as z:
        for f in DATA_FILES:
            with z.open("yaml/%s/%s" % (variant_dir, f)) as src, open(os.path.join(out, f), "wb") as dst:
                shutil.copyfileobj(src, dst)
    return out


def code_root(root):
    """The sys.path entry under which ``src.*`` of a reference root imports: the root itself for a source tree, the bytecode
    archive for a staged one."""
    arc = os.path.join(root, ARCHIVE)
    return arc if os.path.isfile(arc) else root


if __name__ == "__main__":
    m = build(verbose=True)
    print(json.dumps(m) if m else "reference tree not found at %s: nothing built" % SRC_ROOT)
