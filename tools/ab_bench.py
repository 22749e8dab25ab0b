"""Same-box A/B of two builds of libvpr_b200.so (how the round-2 numbers that say "same box, alternating" were taken).

Boxes of the pool differ by ~3 % on the headline step, so two revisions are only comparable inside ONE gpurun call.
This tool builds the library of two git revisions (or the working tree: "WORKTREE") here, on the CPU box, into
`_ab_<name>.so` at the repository root (git-ignored, shipped with the gpurun snapshot), and prints the command that swaps
them in place on the GPU box and runs a short bench with each, twice, alternating:

    python tools/ab_bench.py HEAD~1 WORKTREE            # build both, print the gpurun command
    python tools/ab_bench.py HEAD~1 WORKTREE --knn      # include the kNN leg in each run

Each revision is built in a temporary `git worktree` with its own sources AND its own build stamp, so a stale library
can never stand in for one of the two (that happened once while the stamp file was still tracked by git: a reverted
stamp matched the reverted sources and the build was skipped).  Remove the `_ab_*.so` files afterwards.
"""
import argparse
import importlib.util
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "visual-place-recognition-and-geopose-estimation_b200"


def build_tree(tree: str) -> str:
    spec = importlib.util.spec_from_file_location("_ab_build", os.path.join(tree, PKG, "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=True)


def build_rev(rev: str, out: str) -> None:
    if rev == "WORKTREE":
        shutil.copyfile(build_tree(ROOT), out)
        return
    tmp = tempfile.mkdtemp(prefix="vpr_ab_")
    try:
        subprocess.run(["git", "-C", ROOT, "worktree", "add", "-q", "--detach", tmp, rev], check=True)
        shutil.copyfile(build_tree(tmp), out)
    finally:
        subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", tmp], check=False)
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rev_a")
    ap.add_argument("rev_b")
    ap.add_argument("--knn", action="store_true", help="keep the kNN leg in the bench runs")
    a = ap.parse_args()
    names = []
    for tag, rev in (("a", a.rev_a), ("b", a.rev_b)):
        out = os.path.join(ROOT, f"_ab_{tag}.so")
        sys.stderr.write(f"building {rev} -> {os.path.basename(out)}\n")
        build_rev(rev, out)
        names.append((tag, rev, os.path.basename(out)))
    flags = "--no-config4 --no-strong --no-kernel-table" + ("" if a.knn else " --no-knn")
    show = ('python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); '
            'print(sys.argv[1], d[\\"ms_per_step\\"], d[\\"finetune_step\\"][\\"ms_per_step\\"], d[\\"roofline\\"][\\"kernel_ms\\"]'
            + (', d[\\"knn\\"][\\"value\\"]' if a.knn else "") + ')"')
    body = "; ".join(f"cp {so} $L; timeout 150 python bench.py {flags} 2>/dev/null | {show} '{rev}'" for _, rev, so in names)
    print(f"gpurun --timeout 600 -- 'L={PKG}/libvpr_b200.so; for i in 1 2; do {body}; done'")
    print("# afterwards: rm _ab_a.so _ab_b.so; python -c \"import __graft_entry__ as g; g.build()\"   (the working tree's own library)")


if __name__ == "__main__":
    main()
