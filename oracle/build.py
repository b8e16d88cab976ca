"""TEST INFRASTRUCTURE ONLY — builds oracle/c/*.c into oracle/_build/liboracle.so (gcc, OpenMP)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "liboracle.so")
SRCS = [os.path.join(HERE, "c", "knn_oracle.c")]


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SRCS):
        return OUT
    # -ffp-contract=off: every rounding in the oracle is explicit (fmaf where fused, plain ops elsewhere)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-o", OUT] + SRCS + ["-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
