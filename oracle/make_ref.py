"""Stages the UNMODIFIED reference package into the git-ignored `oracle/_ref/` -- TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

Why: the reference (lukebhan/PDEControlGym) is pure Python + numpy, so "building" it is a file copy.  Its `setup.py`
declares no packages (pip would install nothing) and `pde_control_gym/__init__.py` does not parse, so the
files are copied verbatim and imported through `oracle/ref_loader.py`, which bypasses the broken `__init__`.
`oracle/_ref/` is listed in `.gitignore` (no reference source ever enters the history) but not in
`.gpurunignore`, so the staged copy travels to the GPU box like the built `.so` files do: there it is what
`bench.py --impl reference` / `cpu_baseline` time (`kind: "reference"`) and what `tests/test_oracle_vs_reference.py`
pins the numpy oracle to.  Nothing under `pdecontrolgym_b200/` may import it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PDECG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
# the package (168 KB) + the one data file the reference's NavierStokes examples read (target.npz: U_ref, K5)
ITEMS = ["pde_control_gym", os.path.join("examples", "NavierStokes", "target.npz"),
         os.path.join("examples", "NavierStokes", "NS2Doptimization.py")]


def _same_tree(a, b):
    if os.path.isfile(a):
        return os.path.isfile(b) and filecmp.cmp(a, b, shallow=False)
    c = filecmp.dircmp(a, b, ignore=["__pycache__"])
    if c.left_only or c.right_only or c.diff_files or c.funny_files:
        return False
    return all(_same_tree(os.path.join(a, d), os.path.join(b, d)) for d in c.common_dirs)


def stage(verbose=True) -> bool:
    """Copy the reference files into oracle/_ref (idempotent).  Returns False when the source tree is absent."""
    if not os.path.isdir(os.path.join(SRC, "pde_control_gym", "src")):
        if verbose:
            print(f"[make_ref] {SRC} not present: keeping whatever is staged in {DST}")
        return False
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if not os.path.exists(s):
            continue
        if os.path.exists(d) and _same_tree(s, d):
            continue
        if os.path.isdir(d):
            shutil.rmtree(d)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            shutil.copy2(s, d)
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"verbatim copy of {SRC} ({', '.join(ITEMS)}) made by oracle/make_ref.py; not tracked by git\n")
    if verbose:
        print(f"[make_ref] staged the reference into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() or os.path.isdir(os.path.join(DST, "pde_control_gym", "src")) else 1)
