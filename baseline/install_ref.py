"""Install the UNMODIFIED reference under baseline/_ref (git-ignored, travels to the GPU box with gpurun).

Recipe of SURVEY.md §8(c): a copy of the reference checkout with its three cffi extensions built in place from the
copy's root (the build scripts name their C sources with relative paths). Nothing of the reference is committed; the
copy is only what `bench.py --impl reference` and `tools/bench_reference_configs.py` time.

    python baseline/install_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SOURCE = os.environ.get("JELLYFYSH_REFERENCE", "/root/reference")
REF_ROOT = os.path.join(HERE, "_ref")
BUILD_SCRIPTS = (
    "jellyfysh/potential/merged_image_coulomb_potential/merged_image_coulomb_potential_build.py",
    "jellyfysh/potential/inverse_power_coulomb_bounding_potential/inverse_power_coulomb_bounding_potential_build.py",
    "jellyfysh/scheduler/heap_scheduler/heap_build.py",
)


def installed():
    if not os.path.exists(os.path.join(REF_ROOT, "jellyfysh", "run.py")):
        return False
    for script in BUILD_SCRIPTS:
        directory = os.path.join(REF_ROOT, os.path.dirname(script))
        if not any(name.endswith(".so") for name in os.listdir(directory)):
            return False
    return True


def install(force=False):
    """Returns True when baseline/_ref is usable afterwards."""
    if installed() and not force:
        return True
    if not os.path.exists(os.path.join(REF_SOURCE, "jellyfysh", "run.py")):
        return False
    if os.path.exists(REF_ROOT):
        shutil.rmtree(REF_ROOT)
    shutil.copytree(REF_SOURCE, REF_ROOT, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    for script in BUILD_SCRIPTS:
        subprocess.run([sys.executable, script], cwd=REF_ROOT, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return installed()


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref", "installed" if ok else f"not installed ({REF_SOURCE} absent)")
    sys.exit(0 if ok else 1)
