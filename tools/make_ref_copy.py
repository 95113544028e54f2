"""Copy the unmodified reference (jviquerat/lbm: the lbm/ package and start.py) into the git-ignored
baseline/_ref/ so that it travels to the GPU box with the gpurun snapshot (like the built .so files).

    python tools/make_ref_copy.py            (build container only: needs /root/reference)

Used there by (i) the -m gpu tests that drive the reference's own run() and app classes through
lbm_b200.lattice, (ii) bench.py's Numba CPU baseline (cpu_baseline.kind = "reference") and the
--impl reference arm.  Nothing is modified; results (lbm/save) and caches are left out.  The copy is
never committed (.gitignore: baseline/_ref/)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(os.path.join(SRC, "lbm", "src", "core")):
        print("no reference at %s: nothing copied" % SRC)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(os.path.join(SRC, "lbm"), os.path.join(DST, "lbm"),
                    ignore=shutil.ignore_patterns("save", "__pycache__", "*.pyc"))
    shutil.copy2(os.path.join(SRC, "start.py"), os.path.join(DST, "start.py"))
    n = sum(len(f) for _, _, f in os.walk(DST))
    print("copied %d files to %s" % (n, DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
