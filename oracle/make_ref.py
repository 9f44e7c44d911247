"""Recipe for oracle/_ref/: the reference's own PyTorch modules of the encode path, COPIED (not edited) from
/root/reference/RQ-VAE/models/*.py into the git-ignored oracle/_ref/models/ so that bench.py's reference arm and
`cpu_baseline` can time the REAL reference (`RQVAE.get_indices`, rqvae.py:67-71) on the GPU box's host cores — where
/root/reference does not exist.  TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under oracle/_ref/ is product code and
nothing of it enters the git history (`.gitignore`); it travels to the GPU box with the working tree like the built .so
files.  Called by __graft_entry__.build() when /root/reference is present; a no-op otherwise (bench.py then falls back to
the oracle port and says so).
"""
import os
import shutil

SRC = "/root/reference/RQ-VAE/models"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "models")
FILES = ("__init__.py", "layers.py", "vq.py", "rq.py", "rqvae.py")


def make_ref() -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        src = os.path.join(SRC, f)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(DST, f))
        elif f == "__init__.py":
            open(os.path.join(DST, f), "a").close()
    return True


def ref_available() -> bool:
    return os.path.exists(os.path.join(DST, "rqvae.py"))


if __name__ == "__main__":
    print("oracle/_ref ready:", make_ref())
