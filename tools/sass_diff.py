"""Proves that a change left the shipped kernels untouched: compiles csrc/*.cu of a git revision and of the working tree
with the build flags and compares the SASS opcode sequence of every kernel (addresses, encodings and operands stripped:
ptxas allocates uniform registers differently from build to build of the same source).
    python tools/sass_diff.py <git-ref>
Used at the end of round 1 (no GPU left): every kernel on the default path is instruction-identical to the last revision
that ran on a B200; only opt-in variants / new kernels differ (profiles/r01_sass_identity.md)."""
import os, re, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def kernels(src, incs, out):
    cmd = ["nvcc"] + FLAGS + sum((["-I", i] for i in incs), []) + ["-c", src, "-o", out]
    subprocess.run(cmd, check=True, capture_output=True)
    txt = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    fn, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); fn[cur] = []
            continue
        if cur:
            t = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
            if t:
                # ptxas is not deterministic for the large kernels (two builds of the SAME source differ in uniform-register
                # allocation), so the comparison is on the instruction sequence: opcode + modifiers, operands dropped
                fn[cur].append(t.split()[0] if not t.startswith("@") else " ".join(t.split()[:2]))
    return fn


def main(ref):
    tmp = tempfile.mkdtemp(prefix="sassdiff_")
    subprocess.run(f"git -C {ROOT} archive {ref} parsenet-codebase_b200/csrc include | tar -x -C {tmp}", shell=True, check=True)
    old, new = os.path.join(tmp, "parsenet-codebase_b200/csrc"), os.path.join(ROOT, "parsenet-codebase_b200/csrc")

    def one(f):
        a = kernels(os.path.join(old, f), [old, os.path.join(tmp, "include")], os.path.join(tmp, "old_" + f + ".o"))
        b = kernels(os.path.join(new, f), [new, os.path.join(ROOT, "include")], os.path.join(tmp, "new_" + f + ".o"))
        diff = []
        for name, body in a.items():
            cands = [name, name.replace("iEEv", "iLb0EEEv").replace("xEEv", "xLb0EEEv")]     # knn: added bool template arg
            tgt = next((c for c in cands if c in b), None)
            if tgt is None:
                diff.append((name, "missing"))
            elif body != b[tgt]:
                diff.append((name, f"differs ({len(body)} vs {len(b[tgt])} instructions)"))
        added = [n for n in b if n not in a and n.replace("Lb0E", "") not in a]
        return f, len(a), diff, added

    files = sorted(x for x in os.listdir(old) if x.endswith(".cu"))
    print(f"| file | kernels at {ref} | changed | new kernels |\n|---|---:|---|---:|")
    with ThreadPoolExecutor(8) as ex:
        for f, n, diff, added in ex.map(one, files):
            print(f"| `{f}` | {n} | {'none' if not diff else '; '.join(f'`{a[:60]}` {b}' for a, b in diff)} | {len(added)} |")
    for f in sorted(x for x in os.listdir(new) if x.endswith(".cu") and x not in files):
        print(f"| `{f}` | - | new file | - |")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "HEAD")
