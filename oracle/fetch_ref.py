"""Make the UNMODIFIED reference modules of the hot path travel to the GPU box (test / bench infrastructure only).

The reference is pure Python: there is nothing to compile into oracle/_ref. What bench.py's reference arm and the
same-box cuDNN comparison need are the four module files of SURVEY.md section 8(a) —
    models/unet.py  models/consistency_reglur_memory.py  models/aux_path_memory.py  losses/losses.py
— byte for byte. /root/reference does not exist on the GPU box, so this script copies them into the git-ignored (NOT
gpurun-ignored) directory baseline/_ref/, exactly as BASELINE.md section 4 and SURVEY.md section 8(c) prescribe.
Nothing under baseline/_ref is ever committed, and nothing under pacingpseudo_b200/ imports it.

    python oracle/fetch_ref.py            # in the build container; __graft_entry__.build() calls it too
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PP_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("models/unet.py", "models/consistency_reglur_memory.py", "models/aux_path_memory.py", "losses/losses.py")


def fetch(verbose=False):
    """-> True when baseline/_ref holds the four files (copied now or earlier), False when no reference is around."""
    have = all(os.path.exists(os.path.join(DST, f)) for f in FILES)
    if not os.path.isdir(REF):
        return have
    lines = []
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        lines.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), f))
    with open(os.path.join(DST, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print("\n".join(lines))
    return True


def import_reference():
    """-> (UNet, ConsistencyRegulr, losses module) of the unmodified reference, imported from baseline/_ref.

    Module names `models.*` / `losses.*` are the reference's own (its files import each other by those paths), and the
    drop-in tree (pacingpseudo_b200/dropin) uses the same names on purpose. The drop-in's directories are REGULAR
    packages (they have __init__.py) while the reference's are namespace packages, so sys.path order alone would let
    the drop-in win. The reference files are therefore loaded by explicit file location, registered under their own
    module names for the duration of the import only, with whatever was registered before set aside and restored;
    afterwards they stay reachable as `_pp_ref_.<name>`. A final check refuses to return anything that is not defined
    in a file under baseline/_ref."""
    import importlib.machinery
    import importlib.util
    import types
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        raise FileNotFoundError("baseline/_ref is empty: run `python oracle/fetch_ref.py` where /root/reference exists")

    def clash(k):
        return k in ("models", "losses") or k.startswith("models.") or k.startswith("losses.")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if clash(k)}
    try:
        for pkg in ("models", "losses"):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(DST, pkg)]
            m.__spec__ = importlib.machinery.ModuleSpec(pkg, None, is_package=True)
            m.__spec__.submodule_search_locations = m.__path__
            sys.modules[pkg] = m
        loaded = {}
        for name in ("losses.losses", "models.unet", "models.aux_path_memory", "models.consistency_reglur_memory"):
            path = os.path.join(DST, name.replace(".", os.sep) + ".py")
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            setattr(sys.modules[name.split(".")[0]], name.split(".")[1], mod)
            loaded[name] = mod
    finally:
        mine = {k: sys.modules.pop(k) for k in list(sys.modules) if clash(k)}
        sys.modules.update(saved)
    for k, v in mine.items():   # keep the reference modules reachable under private names
        sys.modules["_pp_ref_." + k] = v
    UNet = loaded["models.unet"].UNet
    ConsistencyRegulr = loaded["models.consistency_reglur_memory"].ConsistencyRegulr
    ref_losses = loaded["losses.losses"]
    for obj in (UNet.forward, ConsistencyRegulr.forward, ref_losses.partial_cross_entropy_loss,
                loaded["models.consistency_reglur_memory"].UNet.forward,
                loaded["models.consistency_reglur_memory"].AuxPath.forward):
        src = os.path.realpath(obj.__code__.co_filename)   # where the code object was compiled from
        if not src.startswith(os.path.realpath(DST) + os.sep):
            raise ImportError("import_reference resolved %r to %s, not to baseline/_ref" % (obj, src))
    return UNet, ConsistencyRegulr, ref_losses


class cpu_only:
    """Context manager for building / running the reference on the HOST cores of a machine that may have a GPU:
    `AuxPath.__init__` calls `.cuda()` on its target tensor (models/aux_path_memory.py:44); inside this block
    `Tensor.cuda` is the identity, so the module stays on the CPU. This is the only shim."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._orig
        return False


if __name__ == "__main__":
    ok = fetch(verbose=True)
    print("baseline/_ref %s" % ("ready" if ok else "NOT available (no reference on this machine)"))
