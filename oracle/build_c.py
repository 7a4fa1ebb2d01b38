"""Build oracle/equss_oracle_c.c (plain C, no dependencies) into oracle/_build/libequss_oracle_c.so.

    python oracle/build_c.py

Test infrastructure, like the rest of oracle/: __graft_entry__.build() compiles it next to the CUDA extension so that
the CPU tests find it; nothing in the product path loads it.  -ffp-contract=off keeps the fp32 arithmetic free of
fused multiply-adds, i.e. the expression trees are the ones written in the source."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "equss_oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libequss_oracle_c.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
               "-Wall", "-Wextra", "-Werror", SRC, "-o", LIB, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for {SRC}:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
