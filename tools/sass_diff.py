"""Which kernels of libfvc_b200.so did a source change really touch?  Compares the SASS of two builds kernel by kernel,
ignoring instruction addresses, constant-bank offsets (they shift when a parameter struct grows) and branch targets.

    python tools/sass_diff.py old/libfvc_b200.so fastvideocodec_b200/libfvc_b200.so [name-filter]

Runs here (cuobjdump, no GPU).  Used in round 2 to keep the ~30 instantiations of k_conv_tc that a change was NOT meant
for byte-identical: ptxas' register allocation of the skip-connection variants is fragile, and a harmless-looking
refactor of shared code (or a new field in the middle of TcParams) cost them 8-14 %."""
import re, subprocess, sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            t = re.sub(r"/\*[0-9a-f]+\*/", "", line).strip()
            t = re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[][]", t)     # parameter offsets
            t = re.sub(r"0x[0-9a-f]{4,}", "ADDR", t)               # branch targets
            d[cur].append(t)
    return d


def main():
    if len(sys.argv) < 3:
        sys.exit(__doc__)
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    flt = sys.argv[3] if len(sys.argv) > 3 else ""
    same = 0
    for k in sorted(set(a) | set(b)):
        if flt and flt not in k:
            continue
        name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
        if k not in a:
            print("new      %6d lines  %s" % (len(b[k]), name))
        elif k not in b:
            print("removed  %6d lines  %s" % (len(a[k]), name))
        elif a[k] != b[k]:
            n = sum(1 for x, y in zip(a[k], b[k]) if x != y) + abs(len(a[k]) - len(b[k]))
            print("changed  %6d -> %6d lines (%d differ)  %s" % (len(a[k]), len(b[k]), n, name))
        else:
            same += 1
    print("%d kernels identical" % same)


if __name__ == "__main__":
    main()
