"""ORACLE (test infrastructure): compile oracle/noise_oracle.c into oracle/_build/libnoise_oracle.so with gcc."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "noise_oracle.c"
OUT = HERE / "_build" / "libnoise_oracle.so"


def build(force: bool = False) -> Path:
    if OUT.exists() and not force and OUT.stat().st_mtime >= SRC.stat().st_mtime:
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", str(OUT), str(SRC), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
