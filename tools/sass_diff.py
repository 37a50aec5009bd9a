"""Which device kernels changed since a git revision?   python tools/sass_diff.py <rev>

Builds amazon-dsstne_b200/csrc of <rev> in a temporary directory with the flags of the Makefile and compares, kernel by
kernel, the SASS instruction text (addresses and encodings stripped) with the objects of the working tree
(amazon-dsstne_b200/build/*.o -- run make first).  Used when code has to change without a GPU at hand: a kernel whose
instruction stream is identical to the one the last green GPU run used cannot have changed behaviour."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def kernels(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, name, body = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = hashlib.md5("\n".join(body).encode()).hexdigest()
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and name:
            body.append(m.group(1))
    if name:
        out[name] = hashlib.md5("\n".join(body).encode()).hexdigest()
    return out


def main():
    rev = sys.argv[1]
    tmp = tempfile.mkdtemp(prefix="sassdiff_")
    subprocess.run(f"git -C {ROOT} archive {rev} amazon-dsstne_b200/csrc include | tar -x -C {tmp}", shell=True, check=True)
    src = os.path.join(tmp, "amazon-dsstne_b200", "csrc")
    procs = []
    for f in sorted(os.listdir(src)):
        if f.endswith(".cu"):
            procs.append((f[:-3], subprocess.Popen(["nvcc", *FLAGS, "-c", f, "-o", os.path.join(tmp, f[:-3] + ".o")], cwd=src,
                                                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)))
    changed = 0
    for tu, p in procs:
        p.wait()
        new_obj = os.path.join(ROOT, "amazon-dsstne_b200", "build", tu + ".o")
        if p.returncode or not os.path.exists(new_obj):
            print(f"{tu}: cannot compare (rev build rc={p.returncode}, working-tree object {'present' if os.path.exists(new_obj) else 'missing'})")
            continue
        a, b = kernels(os.path.join(tmp, tu + ".o")), kernels(new_obj)
        same = [k for k in a if k in b and a[k] == b[k]]
        diff = [k for k in a if k in b and a[k] != b[k]]
        gone = [k for k in a if k not in b]
        new = [k for k in b if k not in a]
        # a kernel that only changed its name (e.g. a new defaulted template parameter) keeps its instruction stream
        renamed = [(k, n) for k in gone for n in new if a[k] == b[n]]
        for k, n in renamed:
            if k in gone and n in new:
                gone.remove(k); new.remove(n); same.append(k)
                print(f"   renamed, identical body: {k} -> {n}")
        changed += len(diff) + len(gone)
        print(f"{tu}: {len(same)} identical, {len(diff)} changed, {len(gone)} removed or renamed, {len(new)} new")
        for k in diff:
            print("   CHANGED", k)
        for k in gone:
            print("   REMOVED", k)
        for k in new:
            print("   new    ", k)
    sys.exit(1 if changed else 0)


if __name__ == "__main__":
    main()
