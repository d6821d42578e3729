"""Device-code identity check used when kernels are added without a GPU at hand: builds the library of a base commit next to the
working tree's and compares the SASS of every kernel both contain.

    python profiles/sass_identity.py <base commit>      # e.g. the last commit whose GPU tests and bench were seen green

Prints the number of kernels in either build, how many of the common ones differ (names listed), and how many are new."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
        elif cur is not None:
            t = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line)      # instruction encodings
            t = re.sub(r"/\*[0-9a-f]{4}\*/", "", t).strip()   # addresses
            if t:
                table[cur].append(t)
    return table


def main():
    base = sys.argv[1]
    with tempfile.TemporaryDirectory() as d:
        tar = subprocess.run(["git", "-C", ROOT, "archive", base, "platipy_b200/csrc", "include"], capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", d], input=tar, check=True)
        lib = os.path.join(d, "base.so")
        subprocess.run(["make", "-C", os.path.join(d, "platipy_b200", "csrc"), "OUT=" + lib], check=True, capture_output=True)
        a = kernels(lib)
    subprocess.run(["make", "-C", os.path.join(ROOT, "platipy_b200", "csrc")], check=True, capture_output=True)
    b = kernels(os.path.join(ROOT, "platipy_b200", "libb200reg.so"))
    common = sorted(set(a) & set(b))
    changed = [k for k in common if a[k] != b[k]]
    print(f"base {base}: {len(a)} kernels; working tree: {len(b)}; common {len(common)}, changed {len(changed)}, new {len(set(b) - set(a))}, "
          f"removed {len(set(a) - set(b))}")
    for k in changed:
        print("  changed:", k)
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
