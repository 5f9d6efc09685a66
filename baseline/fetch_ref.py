"""Stage the UNMODIFIED reference for the GPU box: baseline/_ref/ (git-ignored; it travels with the gpurun snapshot
like the built .so files, /root/reference itself does not exist there).

    python baseline/fetch_ref.py            # in the build container, where /root/reference is mounted

1. The contract's install command is tried first:
       python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
   dontLoveBugs/CSPN_monodepth is not a Python distribution (no setup.py / pyproject.toml), so pip refuses it; the outcome is
   recorded in baseline/_ref/INSTALL.txt and in DESIGN.md.
2. The package tree the hot path and its callers live in (`network/`: the two UNets, CSPN_new.py, CSPN_ours.py, pac.py,
   ...; Python sources only) is then copied VERBATIM into baseline/_ref/network so that
   `sys.path.insert(0, "baseline/_ref"); from network.libs.post_process import CSPN_new` imports the reference's own code.
Nothing under baseline/_ref is tracked, compiled into, or imported by the product (cspn_monodepth_b200/): only
`bench.py --impl reference`, the `cpu_baseline` / reference-on-GPU context rows and the drop-in test use it - as the thing
compared against, never as the thing shipped.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("CSPN_REFERENCE_ROOT", "/root/reference")


def main():
    if not os.path.isdir(os.path.join(SRC, "network")):
        print(f"{SRC}/network not found: nothing staged (this is expected on the GPU box)")
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
           "--target", DST, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
    with open(os.path.join(DST, "INSTALL.txt"), "w") as f:
        f.write("$ " + " ".join(cmd) + f"\nexit code {r.returncode}\n" + "\n".join(tail) + "\n")
        f.write("\nThe reference is not a pip-installable distribution; its `network` package was staged verbatim instead "
                "(baseline/fetch_ref.py).\n")
    n = 0
    for root, dirs, files in os.walk(os.path.join(SRC, "network")):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "inplace_abn")]        # in-place ABN: stale binaries, unrelated to the path
        rel = os.path.relpath(root, SRC)
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for name in files:
            if name.endswith(".py"):
                shutil.copy2(os.path.join(root, name), os.path.join(DST, rel, name))
                n += 1
    print(f"pip install exit code {r.returncode} ({tail[-1] if tail else ''}); staged {n} reference source files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
